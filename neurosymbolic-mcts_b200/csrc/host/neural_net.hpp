// neural_net.hpp — C++ mirror of the reference's evaluator interface over the C ABI:
// `NeuralNetPolicy` (src/neural_net.rs:86-363) and `InferenceServer`
// (src/mcts/inference_server.rs:14-140).  Same method names, argument meaning and error
// behaviour (failures answer "no result" per item and never throw); the compute is
// libcaissa_b200's CUDA path — there is no CPU fallback behind these classes.
#pragma once
#include <condition_variable>
#include <deque>
#include <future>
#include <mutex>
#include <optional>
#include <string>
#include <thread>
#include <utility>
#include <vector>

#include "../../../include/caissa_b200.h"

namespace caissa {

struct TimingAccumulator {  // src/mcts/tactical_mcts.rs:137-175
    uint64_t count = 0;
    double total_ms = 0.0;
    void record(double ms) { ++count; total_ms += ms; }
    double mean_ms() const { return count ? total_ms / (double)count : 0.0; }
};

struct Prediction {
    std::vector<float> policy;  // 4672 probabilities (exp already applied)
    float v_logit = 0.0f;       // raw value logit (not tanh'd)
    float k = 0.0f;             // position-independent confidence scalar
};

typedef uint16_t Move;  // from | to << 6 | promo << 12

class NeuralNetPolicy {
  public:
    explicit NeuralNetPolicy(int num_blocks = 6, int hidden = 128, int dtype = CB_DTYPE_BF16, int device = 0);
    ~NeuralNetPolicy();
    NeuralNetPolicy(NeuralNetPolicy&& o) noexcept;
    NeuralNetPolicy(const NeuralNetPolicy&) = delete;

    // Result<(), String>: empty string = Ok.  `path` is a flat little-endian f32 weight blob.
    std::string load(const std::string& path);
    std::string load_blob(const float* blob, size_t n_floats);
    bool is_available() const;

    std::optional<Prediction> predict(const cb_board& board, bool qsearch_completed, float q_result);
    std::vector<std::optional<Prediction>> predict_batch(const std::vector<cb_board>& boards,
                                                         const std::vector<bool>& qsearch_flags,
                                                         const std::vector<float>& q_results);
    // src/tensor.rs:184-213 on the host (gather, sequential f32 sum, divide, uniform fallback)
    std::vector<std::pair<Move, float>> policy_to_move_priors(const std::vector<float>& policy,
                                                              const std::vector<Move>& moves, const cb_board& board) const;
    void print_inference_timing() const;

    TimingAccumulator transfer_to_gpu_timing, forward_timing, transfer_from_gpu_timing;
    cb_handle* handle() { return h_; }

  private:
    void absorb_timing();
    cb_handle* h_ = nullptr;
    int blocks_, hidden_, dtype_, device_;
    bool warned_ = false;
};

class InferenceServer {
  public:
    typedef std::optional<Prediction> Reply;
    InferenceServer(NeuralNetPolicy nn, size_t batch_size);       // ::new
    static InferenceServer* new_mock_biased(float v_logit);        // ::new_mock_biased
    static InferenceServer* new_mock_uniform() { return new_mock_biased(0.0f); }
    ~InferenceServer();
    // request -> future of (policy, v_logit, k) or nullopt           ::predict_async
    std::future<Reply> predict_async(const cb_board& board, bool qsearch_completed, float q_result);

  private:
    struct Request {
        cb_board board;
        bool flag;
        float q;
        std::promise<Reply> reply;
    };
    InferenceServer() {}
    void worker(size_t batch_size);
    std::unique_ptr<NeuralNetPolicy> nn_;
    bool mock_ = false;
    float mock_v_ = 0.0f;
    std::mutex mu_;
    std::condition_variable cv_;
    std::deque<Request> queue_;
    bool closed_ = false;
    std::thread thread_;
};

}  // namespace caissa

// neural_net.cpp — see neural_net.hpp.
#include "neural_net.hpp"

#include <stdio.h>

#include <chrono>
#include <memory>

#include "chess.hpp"

namespace caissa {

NeuralNetPolicy::NeuralNetPolicy(int num_blocks, int hidden, int dtype, int device)
    : blocks_(num_blocks), hidden_(hidden), dtype_(dtype), device_(device) {}

NeuralNetPolicy::~NeuralNetPolicy() {
    if (h_) cb_destroy(h_);
}

NeuralNetPolicy::NeuralNetPolicy(NeuralNetPolicy&& o) noexcept
    : transfer_to_gpu_timing(o.transfer_to_gpu_timing), forward_timing(o.forward_timing),
      transfer_from_gpu_timing(o.transfer_from_gpu_timing), h_(o.h_), blocks_(o.blocks_), hidden_(o.hidden_),
      dtype_(o.dtype_), device_(o.device_), warned_(o.warned_) {
    o.h_ = nullptr;
}

std::string NeuralNetPolicy::load_blob(const float* blob, size_t n_floats) {
    if (h_) { cb_destroy(h_); h_ = nullptr; }
    int rc = cb_create(&h_, device_, blocks_, hidden_, dtype_);
    if (rc != CB_OK) { h_ = nullptr; return "Failed to load model: no CUDA device / unsupported shape (code " + std::to_string(rc) + ")"; }
    rc = cb_load_weights(h_, blob, n_floats);
    if (rc != CB_OK) {
        std::string e = std::string("Failed to load model: ") + cb_last_error(h_);
        cb_destroy(h_);
        h_ = nullptr;
        return e;
    }
    return std::string();
}

std::string NeuralNetPolicy::load(const std::string& path) {
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) return "Model file not found: " + path;
    fseek(f, 0, SEEK_END);
    const long bytes = ftell(f);
    fseek(f, 0, SEEK_SET);
    std::vector<float> blob((size_t)bytes / 4);
    const size_t got = fread(blob.data(), 4, blob.size(), f);
    fclose(f);
    if (got != blob.size()) return "Failed to load model: short read";
    return load_blob(blob.data(), blob.size());
}

bool NeuralNetPolicy::is_available() const { return h_ && cb_is_available(h_); }

void NeuralNetPolicy::absorb_timing() {
    cb_timing t;
    if (cb_get_timing(h_, &t) == CB_OK && t.calls) {
        transfer_to_gpu_timing.record(t.h2d_ms);
        forward_timing.record(t.forward_ms);
        transfer_from_gpu_timing.record(t.d2h_ms);
        cb_reset_timing(h_);
    }
}

std::vector<std::optional<Prediction>> NeuralNetPolicy::predict_batch(const std::vector<cb_board>& boards,
                                                                      const std::vector<bool>& qsearch_flags,
                                                                      const std::vector<float>& q_results) {
    const size_t n = boards.size();
    std::vector<std::optional<Prediction>> out(n);
    if (!is_available() || n == 0) return out;  // vec![None; n]
    std::vector<uint8_t> flags(n);
    for (size_t i = 0; i < n; ++i) flags[i] = i < qsearch_flags.size() && qsearch_flags[i];
    std::vector<float> policy(n * CB_POLICY_SIZE), v(n), k(n);
    const int rc = cb_predict_batch(h_, boards.data(), q_results.data(), flags.data(), (int)n, policy.data(), v.data(), k.data());
    if (rc != CB_OK) {
        if (!warned_) {  // once-only, like the reference's [NN-ERROR] line (src/neural_net.rs:295-313)
            fprintf(stderr, "[NN-ERROR] predict_batch failed: %s\n", cb_last_error(h_));
            warned_ = true;
        }
        return out;
    }
    absorb_timing();
    for (size_t i = 0; i < n; ++i) {
        Prediction p;
        p.policy.assign(policy.begin() + i * CB_POLICY_SIZE, policy.begin() + (i + 1) * CB_POLICY_SIZE);
        p.v_logit = v[i];
        p.k = k[i];
        out[i] = std::move(p);
    }
    return out;
}

std::optional<Prediction> NeuralNetPolicy::predict(const cb_board& board, bool qsearch_completed, float q_result) {
    auto r = predict_batch({board}, {qsearch_completed}, {q_result});
    return std::move(r[0]);
}

std::vector<std::pair<Move, float>> NeuralNetPolicy::policy_to_move_priors(const std::vector<float>& policy,
                                                                          const std::vector<Move>& moves,
                                                                          const cb_board& board) const {
    std::vector<std::pair<Move, float>> result;
    result.reserve(moves.size());
    float total = 0.0f;
    for (Move m : moves) {
        const int idx = cbchess::policy_index(m, board.w_to_move != 0);
        float p = 0.0f;
        if (idx >= 0 && (size_t)idx < policy.size()) { p = policy[idx]; total += p; }
        result.emplace_back(m, p);
    }
    if (total > 0.0f) {
        for (auto& r : result) r.second /= total;
    } else {
        const float u = 1.0f / (float)moves.size();
        for (auto& r : result) r.second = u;
    }
    return result;
}

void NeuralNetPolicy::print_inference_timing() const {
    if (!forward_timing.count) return;
    printf("NN inference timing over %llu batches: to-GPU %.3f ms, forward %.3f ms, from-GPU %.3f ms (means)\n",
           (unsigned long long)forward_timing.count, transfer_to_gpu_timing.mean_ms(), forward_timing.mean_ms(),
           transfer_from_gpu_timing.mean_ms());
}

// ------------------------------------------------------------------ InferenceServer
InferenceServer::InferenceServer(NeuralNetPolicy nn, size_t batch_size) {
    nn_.reset(new NeuralNetPolicy(std::move(nn)));
    thread_ = std::thread([this, batch_size] { worker(batch_size); });
}

InferenceServer* InferenceServer::new_mock_biased(float v_logit) {
    InferenceServer* s = new InferenceServer();
    s->mock_ = true;
    s->mock_v_ = v_logit;
    s->thread_ = std::thread([s] { s->worker(1); });
    return s;
}

InferenceServer::~InferenceServer() {
    {
        std::lock_guard<std::mutex> l(mu_);
        closed_ = true;
    }
    cv_.notify_all();
    if (thread_.joinable()) thread_.join();
}

std::future<InferenceServer::Reply> InferenceServer::predict_async(const cb_board& board, bool qsearch_completed,
                                                                   float q_result) {
    Request r;
    r.board = board;
    r.flag = qsearch_completed;
    r.q = q_result;
    std::future<Reply> f = r.reply.get_future();
    {
        std::lock_guard<std::mutex> l(mu_);
        queue_.push_back(std::move(r));
    }
    cv_.notify_one();
    return f;
}

// Block on the first request, then gather more until `batch_size` or 500 us
// (src/mcts/inference_server.rs:35-55), evaluate, answer each request.
void InferenceServer::worker(size_t batch_size) {
    for (;;) {
        std::vector<Request> reqs;
        {
            std::unique_lock<std::mutex> l(mu_);
            cv_.wait(l, [this] { return closed_ || !queue_.empty(); });
            if (queue_.empty()) break;  // closed and drained
            reqs.push_back(std::move(queue_.front()));
            queue_.pop_front();
            const auto deadline = std::chrono::steady_clock::now() + std::chrono::microseconds(500);
            while (reqs.size() < batch_size) {
                if (queue_.empty()) {
                    if (cv_.wait_until(l, deadline, [this] { return closed_ || !queue_.empty(); }) == false) break;
                    if (queue_.empty()) break;
                }
                reqs.push_back(std::move(queue_.front()));
                queue_.pop_front();
            }
        }
        if (mock_) {  // zero policy, fixed v_logit, k = 0.5 (src/mcts/inference_server.rs:101-117)
            for (auto& r : reqs) {
                Prediction p;
                p.policy.assign(CB_POLICY_SIZE, 0.0f);
                p.v_logit = mock_v_;
                p.k = 0.5f;
                r.reply.set_value(std::move(p));
            }
            continue;
        }
        std::vector<cb_board> boards;
        std::vector<bool> flags;
        std::vector<float> q;
        for (auto& r : reqs) { boards.push_back(r.board); flags.push_back(r.flag); q.push_back(r.q); }
        auto res = nn_->predict_batch(boards, flags, q);
        for (size_t i = 0; i < reqs.size(); ++i) reqs[i].reply.set_value(std::move(res[i]));
    }
    if (nn_) nn_->print_inference_timing();
}

}  // namespace caissa

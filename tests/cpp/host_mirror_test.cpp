// Exercises the C++ mirror of the reference's evaluator interface (NeuralNetPolicy /
// InferenceServer).  Mode "mock": CPU-only checks that mirror tests/unit/inference_server_tests.rs
// (reply shape and (policy, v_logit, k) semantics of the mock servers) and the no-GPU error
// behaviour.  Mode "gpu <blob.bin> <boards.bin> <n>": real evaluation through the server from
// many threads, printing v_logit per board for the Python test to compare.
#include <stdio.h>
#include <string.h>

#include <thread>
#include <vector>

#include "../../neurosymbolic-mcts_b200/csrc/host/neural_net.hpp"

using namespace caissa;

static int fail(const char* what) {
    fprintf(stderr, "FAIL: %s\n", what);
    return 1;
}

int main(int argc, char** argv) {
    if (argc >= 2 && !strcmp(argv[1], "mock")) {
        cb_board start;
        cb_chess_startpos(&start);
        std::unique_ptr<InferenceServer> biased(InferenceServer::new_mock_biased(2.0f));
        auto r = biased->predict_async(start, true, 0.0f).get();
        if (!r || r->policy.size() != CB_POLICY_SIZE) return fail("mock reply shape");
        if (r->v_logit != 2.0f || r->k != 0.5f) return fail("mock biased values");
        std::unique_ptr<InferenceServer> uni(InferenceServer::new_mock_uniform());
        std::vector<std::future<InferenceServer::Reply>> fs;
        for (int i = 0; i < 32; ++i) fs.push_back(uni->predict_async(start, i % 2 == 0, (float)i));
        for (auto& f : fs) {
            auto x = f.get();
            if (!x || x->v_logit != 0.0f) return fail("mock uniform values");
        }
        // policy_to_move_priors on the host: uniform fallback on an all-zero policy
        NeuralNetPolicy nn;
        uint16_t mv[256];
        const int n = cb_chess_legal_moves(&start, mv, nullptr);
        if (n != 20) return fail("startpos has 20 legal moves");
        auto pri = nn.policy_to_move_priors(r->policy, std::vector<Move>(mv, mv + n), start);
        for (auto& p : pri)
            if (p.second != 1.0f / 20.0f) return fail("uniform fallback");
        // unloaded evaluator: not available, predict answers "None", load of a missing file is Err(String)
        if (nn.is_available()) return fail("unloaded evaluator reports available");
        if (nn.predict(start, true, 0.0f)) return fail("predict without a model must answer nullopt");
        if (nn.load("/nonexistent/model.bin").rfind("Model file not found", 0) != 0) return fail("load error text");
        printf("mock ok\n");
        return 0;
    }
    if (argc >= 5 && !strcmp(argv[1], "gpu")) {
        const int n = atoi(argv[4]);
        std::vector<cb_board> boards(n);
        FILE* f = fopen(argv[3], "rb");
        if (!f || fread(boards.data(), sizeof(cb_board), n, f) != (size_t)n) return fail("boards file");
        fclose(f);
        NeuralNetPolicy nn(6, 128, CB_DTYPE_BF16, 0);
        const std::string err = nn.load(argv[2]);
        if (!err.empty()) return fail(err.c_str());
        if (!nn.is_available()) return fail("not available after load");
        auto direct = nn.predict_batch(boards, std::vector<bool>(n, true), std::vector<float>(n, 0.25f));
        InferenceServer server(std::move(nn), 16);
        std::vector<float> v(n, 0.0f), psum(n, 0.0f);
        std::vector<std::thread> th;
        for (int i = 0; i < n; ++i)
            th.emplace_back([&, i] {
                auto r = server.predict_async(boards[i], true, 0.25f).get();
                if (r) {
                    v[i] = r->v_logit;
                    for (float p : r->policy) psum[i] += p;
                }
            });
        for (auto& t : th) t.join();
        for (int i = 0; i < n; ++i) {
            if (!direct[i]) return fail("direct predict_batch item is None");
            printf("%d %.8f %.8f %.6f %.6f\n", i, v[i], direct[i]->v_logit, psum[i], direct[i]->k);
        }
        return 0;
    }
    fprintf(stderr, "usage: host_mirror_test mock | gpu <blob> <boards> <n>\n");
    return 2;
}

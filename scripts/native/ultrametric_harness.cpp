// Sanitizer harness for laml_b200/csrc/pmm_ultrametric.cpp (host code of the library): random forests with free /
// fixed / zero branches, cold and warm starts, enough edges for the threaded path.  Build and run:
//   scripts/native/sanitize_ultrametric.sh
#include <cstdint>
#include <cstdio>
#include <random>
#include <vector>

extern "C" int pmm_ultrametric_solve(int32_t n, const int32_t *parent, const int8_t *kind, const double *fixed_len, double lo,
                                     double hi, const double *S, double nu, double eps_nu, const double *H0, double mu_final,
                                     double warm_theta, double warm_mu, double *d_out, double *H_out, int32_t *newton_iters);

int main() {
    std::mt19937_64 rng(7);
    std::uniform_real_distribution<double> U(0.0, 1.0);
    int failures = 0;
    for (int rep = 0; rep < 6; rep++) {
        const int N = rep < 4 ? 5 + rep * 40 : 30000;            // the last two cross the threading threshold
        const int n = 2 * N - 1;
        std::vector<int32_t> parent(n, -1);
        // random binary tree: node ids N.. are internal; join two random open nodes until one is left
        std::vector<int> open(N);
        for (int i = 0; i < N; i++) open[i] = i;
        int next = N;
        while (open.size() > 1) {
            for (int k = 0; k < 2; k++) {
                const size_t j = (size_t)(U(rng) * open.size()) % open.size();
                parent[open[j]] = next;
                open[j] = open.back(); open.pop_back();
            }
            open.push_back(next++);
        }
        std::vector<int8_t> kind(n, 0);
        std::vector<double> fixed(n, 0.0), S((size_t)5 * n), d(n), H(n + 1), H2(n + 1), d2(n);
        for (int v = N; v < n; v++) {
            const double u = U(rng);
            if (parent[v] >= 0 && u < 0.05) { kind[v] = 1; fixed[v] = 0.02 + 0.2 * U(rng); }
            else if (parent[v] >= 0 && u < 0.08) kind[v] = 2;
        }
        for (auto &x : S) x = 1e-6 + U(rng) * 3.0;
        int32_t it = 0, it2 = 0;
        const double nu = rep % 2 ? 0.0 : 0.3;
        const int rc = pmm_ultrametric_solve(n, parent.data(), kind.data(), fixed.data(), 0.005, 10.0, S.data(), nu, 1e-5, nullptr,
                                             1e-11, 1e-6, 1e-9, d.data(), H.data(), &it);
        for (auto &x : S) x *= 0.95 + 0.1 * U(rng);
        const int rc2 = rc == 0 ? pmm_ultrametric_solve(n, parent.data(), kind.data(), fixed.data(), 0.005, 10.0, S.data(), nu, 1e-5,
                                                        H.data(), 1e-11, 1e-6, 1e-9, d2.data(), H2.data(), &it2) : rc;
        std::printf("N=%d rc=%d iters=%d warm rc=%d iters=%d height=%.6f\n", N, rc, it, rc2, it2, rc == 0 ? H[n] : 0.0);
        if (rc < 0 || rc2 < 0 || (rc == 0 && it2 >= it)) failures++;
    }
    return failures;
}

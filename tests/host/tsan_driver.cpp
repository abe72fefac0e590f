// TEST INFRASTRUCTURE: runs the emulated library (built with -fsanitize=thread) through a few RK4 steps so that
// ThreadSanitizer can check the kernels' synchronisation: every shared-memory exchange between the threads of a CTA
// must be ordered by a barrier (std::barrier gives TSan the happens-before edges that __syncthreads / __syncwarp
// give the hardware), and the ranks of a multi-GPU run may only touch each other's buffers across the flag barrier.
//   usage: tsan_driver <N0> <N1> <N2> <precision 0|1> <dealias 0|1|2> <solver 0|1|2> [nranks]
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>
#include <complex>
#include "../../include/sdns_b200.h"

static void check(int rc, const char* what) { if (rc) { fprintf(stderr, "%s: %s\n", what, sdns_last_error()); exit(2); } }

int main(int argc, char** argv) {
    if (argc < 7) return 1;
    const int N[3] = {atoi(argv[1]), atoi(argv[2]), atoi(argv[3])};
    const int prec = atoi(argv[4]), dealias = atoi(argv[5]), solver = atoi(argv[6]);
    const int P = argc > 7 ? atoi(argv[7]) : 1;
    std::vector<sdns_plan*> plans(P);
    std::vector<std::vector<char>> handles(P, std::vector<char>(64));
    std::vector<std::vector<char>> ws(P);
    for (int r = 0; r < P; ++r) {
        sdns_config c; memset(&c, 0, sizeof c);
        c.abi_version = SDNS_ABI_VERSION;
        for (int i = 0; i < 3; ++i) { c.N[i] = N[i]; c.L[i] = 6.283185307179586; c.kcut[i] = -1; }
        c.precision = prec; c.dealias = dealias; c.solver = solver;
        c.convection = solver == SDNS_MHD ? SDNS_CONV_DIVERGENCE : SDNS_CONV_VORTEX;
        c.mask_nyquist = 1; c.rank = r; c.nranks = P;
        check(sdns_plan_create(&plans[r], &c), "plan_create");
        if (P > 1) { check(sdns_comm_alloc(plans[r]), "comm_alloc"); check(sdns_comm_handle(plans[r], handles[r].data()), "comm_handle"); }
        else {
            size_t n = 0; check(sdns_workspace_bytes(plans[r], &n), "workspace_bytes");
            ws[r].assign(n + 512, 0);
            char* b = ws[r].data(); b += (256 - (uintptr_t)b % 256) % 256;
            check(sdns_plan_set_workspace(plans[r], b, n), "set_workspace");
        }
    }
    std::vector<char> all(64 * P);
    for (int r = 0; r < P; ++r) memcpy(all.data() + 64 * r, handles[r].data(), 64);
    const size_t cs = prec ? 16 : 8;
    const size_t nel = (size_t)(solver == SDNS_MHD ? 6 : 3) * N[0] * (N[1] / P) * (N[2] / 2 + 1);
    auto rank_fn = [&](int r) {
        if (P > 1) check(sdns_comm_open(plans[r], all.data(), P), "comm_open");
        std::vector<char> u(nel * cs), u1(nel * cs), u2(nel * cs);
        srand(7 + r);
        for (size_t i = 0; i < nel * 2; ++i) {
            const double v = 1e-2 * (rand() / (double)RAND_MAX - 0.5);
            if (prec) ((double*)u.data())[i] = v; else ((float*)u.data())[i] = (float)v;
        }
        for (int s = 0; s < 2; ++s) check(sdns_rk4_step(plans[r], u.data(), u1.data(), u2.data(), 1e-3, 1e-2, 1e-2, nullptr), "rk4_step");
        std::vector<char> phys((size_t)6 * (3 * N[0] / 2) / P * (3 * N[1] / 2) * (3 * N[2] / 2) * (cs / 2) + 1024);
        // every transition between operations that write into peers: backward -> backward -> rhs -> forward -> forward -> rhs
        check(sdns_backward(plans[r], SDNS_SPACE_T, 3, u.data(), phys.data()), "backward");
        check(sdns_backward(plans[r], SDNS_SPACE_T, 3, u.data(), phys.data()), "backward");
        check(sdns_compute_rhs(plans[r], u1.data(), u.data(), 1e-2, 1e-2, nullptr, nullptr), "compute_rhs");
        check(sdns_forward(plans[r], SDNS_SPACE_T, 3, phys.data(), u.data()), "forward");
        check(sdns_forward(plans[r], SDNS_SPACE_T, 3, phys.data(), u.data()), "forward");
        check(sdns_compute_rhs(plans[r], u1.data(), u.data(), 1e-2, 1e-2, nullptr, nullptr), "compute_rhs");
        check(sdns_backward(plans[r], SDNS_SPACE_TP, 3, u.data(), phys.data()), "backward");
    };
    std::vector<std::thread> th;
    for (int r = 0; r < P; ++r) th.emplace_back(rank_fn, r);
    for (auto& t : th) t.join();
    for (int r = 0; r < P; ++r) sdns_plan_destroy(plans[r]);
    printf("tsan_driver done\n");
    return 0;
}

// bench/mpc_bench.cpp -- ROS-free C++ harness that links the solver core directly.
//
//   mpc_bench latency [calls]      BASELINE config 1: repeated MPC::Solve through the adapter class
//                                  (batch of one, host buffers, H2D/D2H included); p50/p99 latency
//   mpc_bench batch [B] [reps]     config 2: one handle, B problems per call through the C ABI
//   mpc_bench multi G [B] [reps] [max_iter]
//                                  config 3: ONE batch of B (65,536) problems split into G contiguous slices, one host
//                                  thread + handle + packed page-locked buffer per GPU, whole control tick per slice
//                                  (pre-step + solve + post-step); wall time from the first submit to the last gather
//                                  into the caller's arrays (no collective, host gather only)
//   mpc_bench poly c0 c1 .. cn     one MPC::Solve with a path polynomial of order n (FG_eval takes any order,
//                                  mpc_planner.cpp:186-190), state (0, 0, 0, 0.3, c0, -0.1), then a cubic again
// Prints one JSON object per run.
#include "mpc_planner.h"
#include "mpc_b200.h"

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

extern "C" {
int mpcgen_num_waypoints(double path_length);
void mpcgen_problems(uint64_t seed, int batch, double path_length, double *wx, double *wy, double *pose, double *vel,
                     int *kind_out);
}

static double now_s()
{
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

static int run_latency(int calls)
{
    MPC mpc;
    std::map<std::string, double> p;   // mpc_params.yaml values through the LoadParams keys
    p["DT"] = 0.1; p["STEPS"] = 20; p["REF_CTE"] = 0; p["REF_ETHETA"] = 0; p["REF_V"] = 0.5; p["W_CTE"] = 100;
    p["W_EPSI"] = 0; p["W_V"] = 1000; p["W_ANGVEL"] = 100; p["W_A"] = 50; p["W_DANGVEL"] = 0; p["W_DA"] = 0;
    p["ANGVEL"] = 1.5; p["MAXTHR"] = 1.0; p["BOUND"] = 1e3;
    mpc.LoadParams(p);
    // problem #0 of seed 20261019 on the infinity track
    const int M = mpcgen_num_waypoints(5.0);
    std::vector<double> wx(M), wy(M), pose(3), vel(3);
    mpcgen_problems(20261019ULL, 1, 5.0, wx.data(), wy.data(), pose.data(), vel.data(), nullptr);
    mpc_b200_params prm; mpc_b200_params_yaml_default(&prm); prm.delay_mode = 0;
    mpc_b200_handle *h = nullptr;
    if (mpc_b200_create(&h, &prm, 1, 0) != MPC_B200_OK) { fprintf(stderr, "no CUDA device\n"); return 2; }
    double co[4], st6[6];
    mpc_b200_prestep_batch(h, 1, M, wx.data(), wy.data(), pose.data(), vel.data(), co, st6, nullptr);
    mpc_b200_destroy(h);
    Eigen::VectorXd state(6), coeffs(4);
    for (int i = 0; i < 6; i++) state[i] = st6[i];
    for (int i = 0; i < 4; i++) coeffs[i] = co[i];
    for (int i = 0; i < 50; i++) mpc.Solve(state, coeffs);
    std::vector<double> t(calls);
    std::vector<double> r;
    for (int i = 0; i < calls; i++) {
        const double t0 = now_s();
        r = mpc.Solve(state, coeffs);
        t[i] = now_s() - t0;
    }
    std::sort(t.begin(), t.end());
    double mean = 0; for (double x : t) mean += x; mean /= calls;
    printf("{\"mode\": \"latency\", \"calls\": %d, \"p50_us\": %.2f, \"p99_us\": %.2f, \"mean_us\": %.2f, \"max_us\": %.2f, "
           "\"w0\": %.12g, \"a0\": %.12g, \"status\": %d, \"iters\": %d, \"kkt\": %.3e}\n",
           calls, 1e6 * t[calls / 2], 1e6 * t[(size_t)(0.99 * calls)], 1e6 * mean, 1e6 * t[calls - 1], r[0], r[1],
           mpc.last_status(), mpc.last_iterations(), mpc.last_kkt_error());
    return 0;
}

static int run_poly(int n, char **c)
{
    MPC mpc;
    std::map<std::string, double> p;
    p["DT"] = 0.1; p["STEPS"] = 20; p["REF_CTE"] = 0; p["REF_ETHETA"] = 0; p["REF_V"] = 0.5; p["W_CTE"] = 100;
    p["W_EPSI"] = 0; p["W_V"] = 1000; p["W_ANGVEL"] = 100; p["W_A"] = 50; p["W_DANGVEL"] = 0; p["W_DA"] = 0;
    p["ANGVEL"] = 1.5; p["MAXTHR"] = 1.0; p["BOUND"] = 1e3;
    mpc.LoadParams(p);
    Eigen::VectorXd state(6), coeffs(n), cubic(n < 4 ? n : 4);
    for (int i = 0; i < n; i++) coeffs[i] = atof(c[i]);
    for (int i = 0; i < cubic.size(); i++) cubic[i] = coeffs[i];
    state[0] = 0; state[1] = 0; state[2] = 0; state[3] = 0.3; state[4] = coeffs[0]; state[5] = -0.1;
    const std::vector<double> r = mpc.Solve(state, coeffs);
    const int st = mpc.last_status(), it = mpc.last_iterations();
    const std::vector<double> r3 = mpc.Solve(state, cubic);        // the handle is back on cubics
    printf("{\"mode\": \"poly\", \"order\": %d, \"w0\": %.12g, \"a0\": %.12g, \"status\": %d, \"iters\": %d, "
           "\"w0_cubic\": %.12g, \"a0_cubic\": %.12g, \"status_cubic\": %d}\n",
           n - 1, r[0], r[1], st, it, r3[0], r3[1], mpc.last_status());
    return 0;
}

static int run_batch(int B, int reps)
{
    mpc_b200_params prm; mpc_b200_params_yaml_default(&prm); prm.delay_mode = 0;
    mpc_b200_handle *h = nullptr;
    if (mpc_b200_create(&h, &prm, B, 0) != MPC_B200_OK) { fprintf(stderr, "no CUDA device\n"); return 2; }
    if (getenv("MPC_BENCH_TWO_STAGES")) mpc_b200_set_option(h, "narrow_one_stage", 0.0);      // A/B of the latency mode
    const int M = mpcgen_num_waypoints(5.0), N = prm.mpc_steps;
    std::vector<double> wx((size_t)M * B), wy((size_t)M * B), pose(3 * (size_t)B), vel(3 * (size_t)B);
    mpcgen_problems(20261018ULL + 2, B, 5.0, wx.data(), wy.data(), pose.data(), vel.data(), nullptr);
    std::vector<double> co(4 * (size_t)B), st(6 * (size_t)B), u0(2 * (size_t)B), pred(3 * (size_t)N * B), obj(B), kkt(B);
    std::vector<int32_t> status(B), iters(B);
    double best = 1e30, kbest = 1e30;
    for (int r = 0; r < reps + 2; r++) {
        const double t0 = now_s();
        mpc_b200_prestep_batch(h, B, M, wx.data(), wy.data(), pose.data(), vel.data(), co.data(), st.data(), nullptr);
        int rc = mpc_b200_solve_batch(h, B, st.data(), co.data(), nullptr, nullptr, u0.data(), pred.data(), obj.data(),
                                      status.data(), iters.data(), kkt.data(), nullptr, nullptr);
        const double t1 = now_s();
        if (rc != MPC_B200_OK) { fprintf(stderr, "solve failed: %s\n", mpc_b200_strerror(rc)); return 3; }
        if (r >= 2) { best = std::min(best, t1 - t0); kbest = std::min(kbest, mpc_b200_last_kernel_seconds(h)); }
    }
    long conv = 0, its = 0; int itmax = 0;
    for (int i = 0; i < B; i++) { conv += status[i] == 1 && kkt[i] <= 1e-8; its += iters[i]; itmax = std::max(itmax, (int)iters[i]); }
    printf("{\"mode\": \"batch\", \"batch\": %d, \"e2e_ms\": %.4f, \"kernel_ms\": %.4f, \"converged\": %ld, "
           "\"solves_per_s_e2e\": %.1f, \"mean_iters\": %.3f, \"max_iters\": %d}\n",
           B, 1e3 * best, 1e3 * kbest, conv, conv / best, (double)its / B, itmax);
    mpc_b200_destroy(h);
    return 0;
}

// page-locked array of the C ABI (mpc_b200_host_alloc), so that the slices move by DMA straight from / to it
template <class T> struct Pinned {
    T *p; size_t n;
    explicit Pinned(size_t n_) : p((T *)mpc_b200_host_alloc(sizeof(T) * n_)), n(n_) { if (p) memset(p, 0, sizeof(T) * n_); }
    ~Pinned() { mpc_b200_host_free(p); }
    T *data() { return p; }
    T &operator[](size_t i) { return p[i]; }
};

static int run_multi(int G, int B, int reps, int max_iter)
{
    const int ndev = mpc_b200_device_count();
    if (ndev < 1) { fprintf(stderr, "no CUDA device\n"); return 2; }
    if (G < 1) G = 1;
    if (G > ndev) { fprintf(stderr, "asked for %d GPUs, %d visible\n", G, ndev); return 2; }
    mpc_b200_params prm; mpc_b200_params_yaml_default(&prm); prm.delay_mode = 0; prm.max_iter = max_iter;
    const int M = mpcgen_num_waypoints(5.0), N = prm.mpc_steps;
    // the caller's arrays: ONE batch of B problems, SoA, page-locked
    Pinned<double> wx((size_t)M * B), wy((size_t)M * B), pose(3 * (size_t)B), vel(3 * (size_t)B);
    Pinned<double> u0(2 * (size_t)B), pred(3 * (size_t)N * B), cmd(2 * (size_t)B), obj(B), kkt(B);
    Pinned<int32_t> status(B), iters(B);
    if (!wx.p || !wy.p || !pose.p || !vel.p || !u0.p || !pred.p || !cmd.p || !obj.p || !kkt.p || !status.p || !iters.p) {
        fprintf(stderr, "host_alloc failed\n"); return 3;
    }
    std::vector<double> vel0(3 * (size_t)B);
    mpcgen_problems(20261018ULL + 3, B, 5.0, wx.data(), wy.data(), pose.data(), vel0.data(), nullptr);
    // one handle per GPU, contiguous slices
    std::vector<mpc_b200_handle *> hs(G, nullptr);
    std::vector<int> lo(G), cnt(G);
    for (int g = 0; g < G; g++) {
        lo[g] = (int)((long long)B * g / G); cnt[g] = (int)((long long)B * (g + 1) / G) - lo[g];
        if (mpc_b200_create(&hs[g], &prm, cnt[g], g) != MPC_B200_OK) { fprintf(stderr, "create failed on GPU %d\n", g); return 3; }
    }
    std::vector<double> times;
    int rc_all = 0;
    for (int r = 0; r < reps + 2; r++) {
        memcpy(vel.data(), vel0.data(), sizeof(double) * 3 * (size_t)B);      // vel is in / out
        const double t0 = now_s();
        std::vector<std::thread> th;
        std::vector<int> rcs(G, 0);
        for (int g = 0; g < G; g++)
            th.emplace_back([&, g]() {
                int rc = mpc_b200_track_slice_submit(hs[g], B, lo[g], cnt[g], M, wx.data(), wy.data(), pose.data(), vel.data(), nullptr,
                                                     u0.data(), pred.data(), cmd.data(), obj.data(), status.data(), iters.data(), kkt.data());
                if (rc == MPC_B200_OK) rc = mpc_b200_track_wait(hs[g]);
                rcs[g] = rc;
            });
        for (auto &t : th) t.join();
        const double t1 = now_s();
        for (int g = 0; g < G; g++) rc_all |= rcs[g];
        if (r >= 2) times.push_back(t1 - t0);
    }
    if (rc_all) { fprintf(stderr, "a slice failed: %s\n", mpc_b200_strerror(rc_all)); return 3; }
    std::sort(times.begin(), times.end());
    const double med = times[times.size() / 2], best = times[0];
    long conv = 0, its = 0; int itmax = 0, n9 = 0, n2 = 0;
    for (int i = 0; i < B; i++) {
        conv += status[i] == 1 && kkt[i] <= 1e-8; its += iters[i]; itmax = std::max(itmax, (int)iters[i]);
        n9 += status[i] == 9; n2 += status[i] == 2;
    }
    double ksec = 0; for (int g = 0; g < G; g++) ksec = std::max(ksec, mpc_b200_last_kernel_seconds(hs[g]));
    printf("{\"mode\": \"multi\", \"gpus\": %d, \"batch\": %d, \"max_iter\": %d, \"one_shot_ms_median\": %.4f, \"one_shot_ms_best\": %.4f, "
           "\"solve_kernel_ms_slowest_gpu\": %.4f, \"converged\": %ld, \"solves_per_s\": %.1f, \"mean_iters\": %.3f, \"max_iters\": %d, "
           "\"status9\": %d, \"status2\": %d, "
           "\"host\": \"C++: std::thread per GPU, one handle each, contiguous slices of the caller's page-locked SoA arrays "
           "(mpc_b200_track_slice_submit / _wait: strided DMA in and out, results land in place = the gather); wall clock from the "
           "first submit to the last wait\"}\n",
           G, B, max_iter, 1e3 * med, 1e3 * best, 1e3 * ksec, conv, conv / med, (double)its / B, itmax, n9, n2);
    for (int g = 0; g < G; g++) mpc_b200_destroy(hs[g]);
    return 0;
}

int main(int argc, char **argv)
{
    if (argc >= 3 && !strcmp(argv[1], "multi"))
        return run_multi(atoi(argv[2]), argc >= 4 ? atoi(argv[3]) : 65536, argc >= 5 ? atoi(argv[4]) : 5, argc >= 6 ? atoi(argv[5]) : 100);
    if (argc >= 2 && !strcmp(argv[1], "latency")) return run_latency(argc >= 3 ? atoi(argv[2]) : 10000);
    if (argc >= 2 && !strcmp(argv[1], "batch")) return run_batch(argc >= 3 ? atoi(argv[2]) : 4096, argc >= 4 ? atoi(argv[3]) : 20);
    if (argc >= 3 && !strcmp(argv[1], "poly")) return run_poly(argc - 2, argv + 2);
    fprintf(stderr, "usage: mpc_bench latency [calls] | batch [B] [reps] | multi G [B] [reps] [max_iter] | poly c0 c1 ..\n");
    return 1;
}

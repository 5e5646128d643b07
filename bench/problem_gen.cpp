// bench/problem_gen.cpp -- synthetic path-tracking problem generator for the benchmarks
// and parity tests (host harness code; no solver logic).
//
// The reference names its test tracks only in prose (README.md:43, "infinity-shaped,
// epitrochoid, square"); SURVEY.md section 8(d) defines them and the sampling rule used here.
// What happens to a generated problem afterwards is the reference's own pre-step
// (mpc_ros/src/mpc_planner_ros.cpp:365-391 down-sampling, mpc_ros/src/driving_state.cpp:196-256).
//
// Built into bench/libmpc_gen.so (C ABI below) and linked by bench/mpc_bench.cpp.
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <vector>

namespace {

struct Path { std::vector<double> x, y; };

// resample a densely sampled closed polyline at constant arc length ds
Path resample_closed(const std::vector<double> &px, const std::vector<double> &py, double ds)
{
    const size_t n = px.size();
    std::vector<double> s(n + 1, 0.0);
    for (size_t i = 0; i < n; i++) {
        const size_t j = (i + 1) % n;
        s[i + 1] = s[i] + std::hypot(px[j] - px[i], py[j] - py[i]);
    }
    const double L = s[n];
    const size_t m = (size_t)std::floor(L / ds);
    Path out;
    out.x.resize(m); out.y.resize(m);
    size_t seg = 0;
    for (size_t q = 0; q < m; q++) {
        const double t = q * ds;
        while (seg + 1 < n && s[seg + 1] < t) seg++;
        const size_t j = (seg + 1) % n;
        const double len = s[seg + 1] - s[seg];
        const double a = len > 0.0 ? (t - s[seg]) / len : 0.0;
        out.x[q] = px[seg] + a * (px[j] - px[seg]);
        out.y[q] = py[seg] + a * (py[j] - py[seg]);
    }
    return out;
}

Path make_path(int kind, double ds)
{
    std::vector<double> px, py;
    const int dense = 200000;
    if (kind == 0) {
        // infinity: Gerono lemniscate, A = 6 m
        const double A = 6.0;
        for (int i = 0; i < dense; i++) {
            const double t = 2.0 * M_PI * i / dense;
            px.push_back(A * std::sin(t));
            py.push_back(A * std::sin(t) * std::cos(t));
        }
    } else if (kind == 1) {
        // epitrochoid R = 3, r = 1, d = 0.5, scale 1.5 m
        const double R = 3.0, r = 1.0, d = 0.5, sc = 1.5;
        for (int i = 0; i < dense; i++) {
            const double t = 2.0 * M_PI * i / dense;
            px.push_back(sc * ((R + r) * std::cos(t) - d * std::cos((R + r) * t / r)));
            py.push_back(sc * ((R + r) * std::sin(t) - d * std::sin((R + r) * t / r)));
        }
    } else {
        // square, side 10 m, axis aligned, sharp corners; exact multiples of ds along each edge
        const double side = 10.0;
        const int per = (int)std::lround(side / ds);
        Path out;
        for (int e = 0; e < 4; e++)
            for (int i = 0; i < per; i++) {
                const double a = i * ds;
                double x, y;
                if (e == 0) { x = a; y = 0.0; }
                else if (e == 1) { x = side; y = a; }
                else if (e == 2) { x = side - a; y = side; }
                else { x = 0.0; y = side - a; }
                out.x.push_back(x); out.y.push_back(y);
            }
        return out;
    }
    return resample_closed(px, py, ds);
}

struct SplitMix64 {
    uint64_t s;
    explicit SplitMix64(uint64_t seed) : s(seed) {}
    uint64_t next()
    {
        uint64_t z = (s += 0x9E3779B97F4A7C15ULL);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
        return z ^ (z >> 31);
    }
    double uni() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }
    double uni(double a, double b) { return a + (b - a) * uni(); }
};

const Path &path_cache(int kind)
{
    static Path paths[3];
    static bool built[3] = { false, false, false };
    if (!built[kind]) { paths[kind] = make_path(kind, 0.05); built[kind] = true; }
    return paths[kind];
}

}  // namespace

extern "C" {

// Number of points of track `kind` (0 infinity, 1 epitrochoid, 2 square) at ds = 0.05 m.
int mpcgen_path_size(int kind) { return (int)path_cache(kind).x.size(); }
void mpcgen_path_copy(int kind, double *x, double *y)
{
    const Path &p = path_cache(kind);
    for (size_t i = 0; i < p.x.size(); i++) { x[i] = p.x[i]; y[i] = p.y[i]; }
}

// Waypoints kept by the reference's down-sampling of a `path_length` window at spacing 0.05:
// every int(path_length/10/0.05)-th point plus the last (mpc_planner_ros.cpp:365-391).
int mpcgen_num_waypoints(double path_length)
{
    const int win = (int)std::lround(path_length / 0.05);
    const int step = (int)(path_length / 10.0 / 0.05);
    int m = 0;
    for (int i = 0; i < win; i += step) m++;
    return m + 1;
}

// Generates `batch` problems (SoA, problem index fastest):
//   wx, wy  M x batch   down-sampled reference window, global frame
//   pose    3 x batch   px, py, theta
//   vel     3 x batch   v, previous w, previous throttle
//   kind    batch       track id
void mpcgen_problems(uint64_t seed, int batch, double path_length,
                     double *wx, double *wy, double *pose, double *vel, int *kind_out)
{
    SplitMix64 rng(seed);
    const int win = (int)std::lround(path_length / 0.05);
    const int step = (int)(path_length / 10.0 / 0.05);
    const int M = mpcgen_num_waypoints(path_length);
    for (int i = 0; i < batch; i++) {
        const int kind = i % 3;
        const Path &P = path_cache(kind);
        const int n = (int)P.x.size();
        const int start = (int)(rng.uni() * n) % n;
        const double lat = rng.uni(-0.3, 0.3);
        const double dth = rng.uni(-0.4, 0.4);
        const double v = rng.uni(0.0, 0.6);
        const double pw = rng.uni(-0.5, 0.5);
        const double pa = rng.uni(-0.5, 0.5);
        const int nx = (start + 1) % n;
        double tx = P.x[nx] - P.x[start], ty = P.y[nx] - P.y[start];
        const double tl = std::hypot(tx, ty);
        tx /= tl; ty /= tl;
        pose[0 * (size_t)batch + i] = P.x[start] - ty * lat;
        pose[1 * (size_t)batch + i] = P.y[start] + tx * lat;
        pose[2 * (size_t)batch + i] = std::atan2(ty, tx) + dth;
        vel[0 * (size_t)batch + i] = v;
        vel[1 * (size_t)batch + i] = pw;
        vel[2 * (size_t)batch + i] = pa;
        if (kind_out) kind_out[i] = kind;
        int m = 0;
        for (int j = 0; j < win; j += step, m++) {
            const int q = (start + j) % n;
            wx[(size_t)m * batch + i] = P.x[q]; wy[(size_t)m * batch + i] = P.y[q];
        }
        const int q = (start + win - 1) % n;
        wx[(size_t)m * batch + i] = P.x[q]; wy[(size_t)m * batch + i] = P.y[q];
        m++;
        (void)M;
    }
}

}  // extern "C"

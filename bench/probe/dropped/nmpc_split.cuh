// nmpc_split.cuh -- the 4 x 4 Riccati sweeps of the plain variant (nmpc_phases.cuh, riccati_backward4 / riccati_forward4)
// spread over FOUR lanes per problem: lane c = 0..3 of a group owns column c of the symmetric value matrix over
// (x, y, theta, v) -- by symmetry also its row c -- and entry c of every vector; eight problems per warp.
//
// Why: a warp pays for an FP64 instruction per INSTRUCTION, not per active lane (profiles/r1_fp64_probe.md), and the sweeps
// are one serial chain per problem.  With one problem per lane a CTA that holds only a few problems (long horizons: 6 lanes
// at N = 100; a single MPC::Solve: 1 lane) runs ~140 FP64 instructions per stage on a nearly empty warp.  Here every lane
// of the group executes the same ~70 instructions on its own column (uniform code, no divergence); what a lane needs of
// the other columns moves by warp shuffles.  Same arithmetic per entry as the one-lane form up to the order of two sums.
//
// Used by nmpc_solve_kernel when the lanes-per-CTA count is a compile-time value <= 8 (control warp 0 carries all eight
// groups).  Device only: the host emulator (tests/emu) runs the one-lane form of the same recursion.
#pragma once
#include "nmpc_phases.cuh"

#if defined(__CUDACC__)
namespace nmpc {

#define NMPC_FULL 0xffffffffu

__device__ __forceinline__ double sel(bool c, double a, double b) { return c ? a : b; }

// Backward sweep.  `p`: the group's problem lane (valid index even when !active: loads stay in bounds),
// `c`: this lane's column.  Returns 1 if every R~_k was positive definite (same value in the four lanes).
template <class SM>
__device__ __forceinline__ int riccati_backward_split(const Params &prm, const SM &sm, int p, bool active, int c, const HessDiag &hd)
{
    const int N = prm.N;
    const int gb = (threadIdx.x & 31) & ~3;          // first lane of the group
    const bool isx = c == 0, isy = c == 1, ist = c == 2, isv = c == 3, hi = c >= 2, odd = (c & 1) != 0;
    // eps_{N-1} = sum of the defect differences d_e - d_theta (each lane a quarter of the stages)
    double eps = 0.0;
    for (int k = c; k < N - 1; k += 4) eps += sm.at(k, D_E, p) - sm.at(k, D_T, p);
    eps += __shfl_xor_sync(NMPC_FULL, eps, 1);
    eps += __shfl_xor_sync(NMPC_FULL, eps, 2);
    // terminal stage: column c of diag(dx, dy, dt + de, dv); vector (0, 0, q_e + de eps, q_v)
    double P0 = isx ? hd.dx : 0.0, P1 = isy ? hd.dy : 0.0, P2 = ist ? hd.dt_ + hd.de : 0.0, P3 = isv ? hd.dv : 0.0;
    double pc = ist ? fma(hd.de, eps, sm.at(N - 1, W_2, p)) : (isv ? sm.at(N - 1, W_0, p) : 0.0);
    double qc_next = sm.at(N - 1, W_1, p);
    const double gam = hd.dc;
    int ok = 1;
    StageCoef q;
#pragma unroll 1
    for (int k = N - 2; k >= 0; k--) {
        load_coef(sm, k, p, q);
        // ---- R~ and its inverse (every lane: the same numbers): P_tt, P_tv from lane theta, P_vv from lane v
        const double Ptt = __shfl_sync(NMPC_FULL, P2, gb + 2), Ptv = __shfl_sync(NMPC_FULL, P3, gb + 2);
        const double Pvv = __shfl_sync(NMPC_FULL, P3, gb + 3);
        const double Rww = q.rw + Ptt, Rwa = Ptv, Raa = q.ra + Pvv;
        const double det = Rww * Raa - Rwa * Rwa;
        if (!(Rww > 0.0) || !(det > 0.0)) ok = 0;
        const double idet = fast_rcp(det);
        const double i11 = Raa * idet, i12 = -Rwa * idet, i22 = Rww * idet;
        // ---- etheta folded into theta (riccati_backward4)
        const double ek = eps - (q.de - q.dth);
        const double Qtt = q.htt + q.hee, Qtv = q.htv + q.hev;
        const double qt = fma(q.hee, ek, q.qe), qv = fma(q.hev, ek, q.qv);
        const double dc = fma(q.a56, ek, q.dc);
        // ---- row c of W = P A4: entries (c, theta) and (c, v); (c, x) = P0, (c, y) = P1
        const double Wrt = P2 + q.a13 * P0 + q.a23 * P1;
        const double Wrv = P3 + q.a14 * P0 + q.a24 * P1;
        // ---- column c of W: lanes theta and v collect W[r][c] from the row owners (three xor rounds; what a lane sends
        //      is what its partner of the round needs: W[.][theta] to lane theta, W[.][v] to lane v)
        const double sA = odd ? Wrv : Wrt, sB = odd ? Wrt : Wrv;
        const double G1 = __shfl_xor_sync(NMPC_FULL, sB, 1), G2 = __shfl_xor_sync(NMPC_FULL, sA, 2);
        const double G3 = __shfl_xor_sync(NMPC_FULL, sB, 3);
        const double Wc0 = hi ? (ist ? G2 : G3) : P0, Wc1 = hi ? (ist ? G3 : G2) : P1;
        const double Wc2 = hi ? (ist ? Wrt : G1) : P2, Wc3 = hi ? (ist ? G1 : Wrv) : P3;
        // ---- column c of M = A4^T W
        const double Mc2 = Wc2 + q.a13 * Wc0 + q.a23 * Wc1;
        const double Mc3 = Wc3 + q.a14 * Wc0 + q.a24 * Wc1;
        // ---- S~ = rows theta, v of W: this lane's entries; its column of the gains
        const double Sw = Wc2, Sa = Wc3;
        const double Kw = -(i11 * Sw + i12 * Sa), Ka = -(i12 * Sw + i22 * Sa);
        if (active) { sm.at(k, W_0 + c, p) = Kw; sm.at(k, W_4 + c, p) = Ka; }
        const double Kw0 = __shfl_sync(NMPC_FULL, Kw, gb), Kw1 = __shfl_sync(NMPC_FULL, Kw, gb + 1);
        const double Kw2 = __shfl_sync(NMPC_FULL, Kw, gb + 2), Kw3 = __shfl_sync(NMPC_FULL, Kw, gb + 3);
        const double Ka0 = __shfl_sync(NMPC_FULL, Ka, gb), Ka1 = __shfl_sync(NMPC_FULL, Ka, gb + 1);
        const double Ka2 = __shfl_sync(NMPC_FULL, Ka, gb + 2), Ka3 = __shfl_sync(NMPC_FULL, Ka, gb + 3);
        // ---- vector part: entry c of p~ = P d + p; all four entries to everybody
        const double tc = fma(P0, q.dx, fma(P1, q.dy, fma(P2, q.dth, fma(P3, q.dv, pc))));
        const double t0 = __shfl_sync(NMPC_FULL, tc, gb), t1 = __shfl_sync(NMPC_FULL, tc, gb + 1);
        const double t2 = __shfl_sync(NMPC_FULL, tc, gb + 2), t3 = __shfl_sync(NMPC_FULL, tc, gb + 3);
        const double pic = fma(gam, dc, qc_next);
        const double ruw = q.qw + t2, rua = q.qa + t3;
        const double kfw = -(i11 * ruw + i12 * rua), kfa = -(i12 * ruw + i22 * rua);
        if (active && isx) { sm.at(k, W_10, p) = kfw; sm.at(k, W_11, p) = kfa; }
        // ---- column c of Q~ = M + Q + gam a_c a_c^T,  a_c = [a51, -1, a56, a54]
        const double ac = isx ? q.a51 : (isy ? -1.0 : (ist ? q.a56 : q.a54));
        const double gc = gam * ac;
        const double Q0 = fma(gc, q.a51, Wc0) + (isx ? q.hxx : 0.0);
        const double Q1 = (Wc1 - gc) + (isy ? hd.dy : 0.0);
        const double Q2 = fma(gc, q.a56, Mc2) + (ist ? Qtt : (isv ? Qtv : 0.0));
        const double Q3 = fma(gc, q.a54, Mc3) + (isv ? hd.dv : (ist ? Qtv : 0.0));
        // ---- column c of P_k = Q~ + S~^T K  (= row c: S~^T K is symmetric)
        P0 = fma(Sw, Kw0, fma(Sa, Ka0, Q0));
        P1 = fma(Sw, Kw1, fma(Sa, Ka1, Q1));
        P2 = fma(Sw, Kw2, fma(Sa, Ka2, Q2));
        P3 = fma(Sw, Kw3, fma(Sa, Ka3, Q3));
        // ---- entry c of p_k = q_s + A^T p~ + a_c pi_c + S~^T k_ff
        const double al = ist ? q.a13 : (isv ? q.a14 : 0.0), be = ist ? q.a23 : (isv ? q.a24 : 0.0);
        const double qq = ist ? qt : (isv ? qv : 0.0);
        pc = fma(Sw, kfw, fma(Sa, kfa, fma(ac, pic, fma(be, t1, fma(al, t0, qq + tc)))));
        qc_next = q.qc;
        eps = ek;
    }
    return ok;
}

// Forward sweep: lane c carries ds[c] (x, y, theta, v); the new entry is one row of  ds' = A4 ds + B du + d  with the
// gains folded into the rows of theta and v:  n_c = s_c + k0 s_x + k1 s_y + k2 s_t + k3 s_v + off.
// Lane x also forms the cte row, lane theta the etheta entry (theta + eps).  Results as riccati_forward4.
template <class SM>
__device__ __forceinline__ void riccati_forward_split(const Params &prm, const SM &sm, int p, bool active, int c)
{
    const int N = prm.N;
    const int gb = (threadIdx.x & 31) & ~3;
    const bool isx = c == 0, isy = c == 1, ist = c == 2, hi = c >= 2;
    const double idt = prm.idt;
    // per-lane slots of the row coefficients: rows x, y take A entries (columns theta, v only), rows theta, v the gains
    const int sl2 = isx ? A_13 : (isy ? A_23 : (ist ? W_2 : W_6));
    const int sl3 = isx ? A_14 : (isy ? A_24 : (ist ? W_3 : W_7));
    const int sl0 = ist ? W_0 : W_4, sl1 = ist ? W_1 : W_5;            // rows theta, v only
    const int sld = isx ? D_X : (isy ? D_Y : (ist ? D_T : D_V));
    const int slf = ist ? W_10 : W_11;                                   // feed-forward / where du goes
    double s = 0.0, eps = 0.0;
#pragma unroll 2
    for (int k = 0; k < N - 1; k++) {
        const double sx = __shfl_sync(NMPC_FULL, s, gb), sy = __shfl_sync(NMPC_FULL, s, gb + 1);
        const double st = __shfl_sync(NMPC_FULL, s, gb + 2), sv = __shfl_sync(NMPC_FULL, s, gb + 3);
        const double k2 = sm.at(k, sl2, p), k3 = sm.at(k, sl3, p);
        const double k0 = hi ? sm.at(k, sl0, p) : 0.0, k1 = hi ? sm.at(k, sl1, p) : 0.0;
        const double off = sm.at(k, sld, p), kf = hi ? sm.at(k, slf, p) : 0.0;
        const double a51 = sm.at(k, A_51, p), a54 = sm.at(k, A_54, p), a56 = sm.at(k, A_56, p);
        const double dcc = sm.at(k, D_C, p), dth = sm.at(k, D_T, p), de = sm.at(k, D_E, p);
        // (rows theta, v: du = K ds + k_ff, the scaled step; n = s + du + d)
        const double du = fma(k0, sx, k1 * sy) + fma(k2, st, fma(k3, sv, kf));
        const double n = s + du + off;
        const double nc = a51 * sx - sy + a54 * sv + a56 * (st + eps) + dcc;
        eps += de - dth;
        const double ne = n + eps;                       // meaningful on lane theta
        __syncwarp();                                    // every lane has read stage k before anybody overwrites it
        if (active) {
            sm.at(k, sld, p) = n;
            if (isx) sm.at(k, D_C, p) = nc;
            if (ist) sm.at(k, D_E, p) = ne;
            if (hi) sm.at(k, slf, p) = du * idt;
        }
        s = n;
    }
}

}  // namespace nmpc
#endif

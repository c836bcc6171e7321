// env_step_core.cuh — the fused Go2 environment step for ONE env, written for one 32-lane warp.
//
// Replaces LeggedRobot.step + post_physics_step of the reference (legged_gym/envs/base/legged_robot.py:60-142 and
// everything they call; go2_env.py:23-60) and the gym.simulate call inside it (physics spec: DESIGN.md section 3).
//
// Execution model: a sequence of PHASES.  Inside a phase every thread works on its own item (a joint, a leg, a collider,
// a height sample, an observation column ...) and threads only communicate through the per-env shared-memory block
// `WarpSmem` ACROSS phase boundaries (GO2_SYNC()).  No shuffles or ballots are used, so all cross-lane sums run in a fixed
// order (bit-reproducible run to run) and the very same source can be executed thread-by-thread on the host:
//   * nvcc, sm_100a : GO2_WIDE / GO2_LEGS guard a block by the thread's role; GO2_SYNC() = __syncwarp() or a CTA barrier.
//   * g++ (tests/emu): they expand to a loop over the threads of the group, giving a faithful functional emulation of the
//     kernel that the CPU test-suite compares against the oracle without a GPU.  The emulation is test tooling, not a product path.
//
// Roles (filled by init_roles()):
//   WIDE  one warp per env, item = lane: joints j = lane (0..11) · colliders c = lane (0..31) · reported bodies · height samples
//         i = lane + 32 k · observation columns i = lane + 32 k · the scalar per-env bookkeeping on lane 0.
//   LEGS  one thread per (env, leg): the serial 6x6 articulated-body recursions along hip -> thigh -> calf, and (redundantly on the 4
//         leg threads of an env) the base's 6x6 inverse and impulse response, so that consecutive LEGS phases only need a warp sync.
// Three families of thread maps exist (init_roles).  "warp per env": LEGS = lanes 0..3 of the env's own warp (28 lanes idle in the heaviest phases).
// "packed": a CTA of 8 warps owns 8 envs and the 32 (env, leg) items fill warp 0, so the long serial leg code is issued once per
// 8 envs instead of once per env; WIDE <-> LEGS transitions are CTA barriers, LEGS -> LEGS transitions stay inside warp 0.
// "half-warp" (H14, the default): 16 threads per env, each running a WIDE block for two "virtual lanes" (T::NV = 2), and dedicated LEGS warps
// compiled as their own instantiation (T::ROLE): 14 envs per 9-warp CTA at 96 registers, 28 envs resident per SM.
#pragma once
#include <stdint.h>
#include <math.h>
#include "../../include/go2_b200.h"

#if defined(__CUDACC__)
#define GO2_HD __device__ __forceinline__
#else
#define GO2_HD inline
#endif

#if defined(__CUDACC__)
#define GO2_EACH
#define GO2_SYNC() go2_phase_sync(L)
#define GO2_SYNC_WARP() __syncwarp()   /* between two phases of the SAME role: the threads of a role that touch one env always share a warp */
#define GO2_FMUL(a, b) __fmul_rn((a), (b))
#define GO2_FADD(a, b) __fadd_rn((a), (b))
#define GO2_LDG(p) __ldg(p)
#define GO2_UNROLL _Pragma("unroll")
#else
#define GO2_EACH for (int tid_ = 0; tid_ < NT; ++tid_) if (Lane& L = lanes[tid_]; true)
#define GO2_SYNC() do { } while (0)
#define GO2_SYNC_WARP() do { } while (0)
#define GO2_FMUL(a, b) ((a) * (b))
#define GO2_FADD(a, b) ((a) + (b))
#define GO2_LDG(p) (*(p))
#define GO2_UNROLL
#endif
// bind S (the item's env scratch), e (its env id) and lane (the item index within the role) for the block that follows
#define GO2_BIND(slot, idx) if (auto& S = SM[slot]; true) if (const int e = L.e0 + (slot), lane = (idx); (void)e, (void)lane, true)
// WIDE: a thread owns T::NV "virtual lanes" of its env (NV = 1: lane = the thread's lane in the env's warp; NV = 2, the half-warp maps: an env
// is served by 16 threads, thread k runs the block for item k and then for item k + 16).  vh is visible inside the block (per-collider state).
// T::ROLE prunes the other role's code at compile time in kernels whose warps are specialised (0 = both roles, 1 = WIDE only, 2 = LEGS only).
#if defined(__CUDACC__)   /* the kernel's map is a compile-time property */
#define GO2_VLANE(vh) ((T::NV == 2 ? (L.lane & 15) : L.lane) + 16 * (vh))
#define GO2_NV_OK(vh) true
#else                     /* the emulation is compiled once (T::NV = 2) and takes the map from init_roles */
#define GO2_VLANE(vh) ((L.nv == 2 ? (L.lane & 15) : L.lane) + 16 * (vh))
#define GO2_NV_OK(vh) ((vh) < L.nv)
#endif
#define GO2_WIDE GO2_EACH if (T::ROLE != 2 && L.own) GO2_UNROLL for (int vh = 0; vh < T::NV; ++vh) if (GO2_NV_OK(vh)) GO2_BIND(L.w, GO2_VLANE(vh))
// the same for the phases that run ONCE per step (post-physics, store, reset) with the loop over the virtual lanes kept ROLLED (-DGO2_COLD_ROLLED=1).
// That code is cold in the instruction cache every time it runs, so one copy instead of two looked attractive; measured on the B200 (round 2,
// 4096 envs): 13.3 k instead of 14.1 k SASS instructions (most second-lane copies are pruned anyway: the lane ranges are compile-time), 107.3 us
// against 106.6 us unrolled — the rolled loop serialises the two lanes' loads.  Default: unrolled.
#if !defined(GO2_COLD_ROLLED)
#define GO2_COLD_ROLLED 0
#endif
#if defined(__CUDACC__) && GO2_COLD_ROLLED
#define GO2_WIDE_COLD GO2_EACH if (T::ROLE != 2 && L.own) _Pragma("unroll 1") for (int vh = 0; vh < T::NV; ++vh) GO2_BIND(L.w, GO2_VLANE(vh))
#else
#define GO2_WIDE_COLD GO2_WIDE
#endif
// items i = lane, lane + 32, ... < n of a WIDE block as a FIXED-trip unrolled loop with a predicate: the loads of all trips are in flight together
// (a `for (i = lane; i < n; i += 32)` loop has a lane-dependent trip count, is not unrolled, and serialises one memory round trip per trip)
#define GO2_STRIDED(i, n) GO2_UNROLL for (int k_ = 0; k_ < ((n) + 31) / 32; ++k_) if (const int i = lane + 32 * k_; i < (n))
#define GO2_LEGS GO2_EACH if (T::ROLE != 1 && L.leg >= 0) GO2_BIND(L.wl, L.leg)

namespace go2 {

// ------------------------------------------------------------------------------------------------ Philox4x32-10
struct U4 { uint32_t x, y, z, w; };
GO2_HD uint32_t mulhi32(uint32_t a, uint32_t b) {
#if defined(__CUDACC__)
  return __umulhi(a, b);
#else
  return (uint32_t)(((uint64_t)a * b) >> 32);
#endif
}
GO2_HD U4 philox(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t h0 = mulhi32(0xD2511F53u, c0), l0 = 0xD2511F53u * c0;
    uint32_t h1 = mulhi32(0xCD9E8D57u, c2), l1 = 0xCD9E8D57u * c2;
    uint32_t n0 = h1 ^ c1 ^ k0, n2 = h0 ^ c3 ^ k1;
    c0 = n0; c1 = l1; c2 = n2; c3 = l0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  return U4{c0, c1, c2, c3};
}
GO2_HD float u01(uint32_t x) { return (float)(x >> 8) * (1.0f / 16777216.0f); }
GO2_HD uint32_t pick(const U4& r, int k) { return k == 0 ? r.x : (k == 1 ? r.y : (k == 2 ? r.z : r.w)); }
// span * u + lo rounded like torch (separate multiply and add, no FMA contraction)
GO2_HD float affine(float span, float u, float lo) { return GO2_FADD(GO2_FMUL(span, u), lo); }
enum { ST_DELAY = 0, ST_NOISE = 1, ST_PUSH = 2, ST_RESET_DR = 3, ST_RESET_STATE = 4, ST_CMD_CB = 5, ST_CMD_RESET = 6 };

// ------------------------------------------------------------------------------------------------ 3-vectors / 3x3
struct V3 { float x, y, z; };
GO2_HD V3 mk(float x, float y, float z) { V3 a; a.x = x; a.y = y; a.z = z; return a; }
GO2_HD V3 operator+(V3 a, V3 b) { return mk(a.x + b.x, a.y + b.y, a.z + b.z); }
GO2_HD V3 operator-(V3 a, V3 b) { return mk(a.x - b.x, a.y - b.y, a.z - b.z); }
GO2_HD V3 operator*(float s, V3 a) { return mk(s * a.x, s * a.y, s * a.z); }
GO2_HD float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
GO2_HD V3 cross(V3 a, V3 b) { return mk(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
GO2_HD V3 ld3(const float* p) { return mk(p[0], p[1], p[2]); }
GO2_HD void st3(float* p, V3 a) { p[0] = a.x; p[1] = a.y; p[2] = a.z; }
struct M3 { float m[9]; };  // row major
GO2_HD V3 mul(const M3& A, V3 v) {
  return mk(A.m[0] * v.x + A.m[1] * v.y + A.m[2] * v.z, A.m[3] * v.x + A.m[4] * v.y + A.m[5] * v.z, A.m[6] * v.x + A.m[7] * v.y + A.m[8] * v.z);
}
GO2_HD V3 mulT(const M3& A, V3 v) {
  return mk(A.m[0] * v.x + A.m[3] * v.y + A.m[6] * v.z, A.m[1] * v.x + A.m[4] * v.y + A.m[7] * v.z, A.m[2] * v.x + A.m[5] * v.y + A.m[8] * v.z);
}
GO2_HD M3 mul(const M3& A, const M3& B) {
  M3 C;
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) C.m[3 * i + j] = A.m[3 * i] * B.m[j] + A.m[3 * i + 1] * B.m[3 + j] + A.m[3 * i + 2] * B.m[6 + j];
  return C;
}
GO2_HD M3 transpose(const M3& A) {
  M3 C;
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) C.m[3 * i + j] = A.m[3 * j + i];
  return C;
}
GO2_HD M3 add(const M3& A, const M3& B) { M3 C; for (int i = 0; i < 9; ++i) C.m[i] = A.m[i] + B.m[i]; return C; }
GO2_HD M3 sub(const M3& A, const M3& B) { M3 C; for (int i = 0; i < 9; ++i) C.m[i] = A.m[i] - B.m[i]; return C; }
GO2_HD M3 ldm(const float* p) { M3 A; for (int i = 0; i < 9; ++i) A.m[i] = p[i]; return A; }
GO2_HD void stm(float* p, const M3& A) { for (int i = 0; i < 9; ++i) p[i] = A.m[i]; }
// r x M  (skew(r) * M) and M x r (M * skew(r))
GO2_HD M3 skew_mul(V3 r, const M3& M) {
  M3 C;
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    V3 col = cross(r, mk(M.m[j], M.m[3 + j], M.m[6 + j]));
    C.m[j] = col.x; C.m[3 + j] = col.y; C.m[6 + j] = col.z;
  }
  return C;
}
GO2_HD M3 mul_skew(const M3& M, V3 r) {  // M * skew(r): row_i x ... (M rx)_i: = -(r x row_i)^T = row_i x r
  M3 C;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    V3 row = cross(mk(M.m[3 * i], M.m[3 * i + 1], M.m[3 * i + 2]), r);
    C.m[3 * i] = row.x; C.m[3 * i + 1] = row.y; C.m[3 * i + 2] = row.z;
  }
  return C;
}
// coordinate-axis rotation Rpc (child -> parent) applied as Rpc*M, M*Rpc^T etc. (c, s) = cos, sin of the joint angle.
// pair (i, j) = (1,2) for x, (2,0) for y, (0,1) for z:  row_i' = c row_i - s row_j ; row_j' = s row_i + c row_j
template <int AX> struct AxPair { static constexpr int i = (AX + 1) % 3, j = (AX + 2) % 3; };
template <int AX> GO2_HD V3 rot_fwd(float c, float s, V3 v) {  // Rpc v
  float a[3] = {v.x, v.y, v.z};
  constexpr int i = AxPair<AX>::i, j = AxPair<AX>::j;
  float ai = c * a[i] - s * a[j], aj = s * a[i] + c * a[j];
  a[i] = ai; a[j] = aj;
  return mk(a[0], a[1], a[2]);
}
template <int AX> GO2_HD V3 rot_inv(float c, float s, V3 v) { return rot_fwd<AX>(c, -s, v); }  // Rpc^T v = E v
template <int AX> GO2_HD M3 rot_sim_fwd(float c, float s, const M3& M) {  // Rpc M Rpc^T
  M3 T = M;
  constexpr int i = AxPair<AX>::i, j = AxPair<AX>::j;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    float a = T.m[3 * i + k], b = T.m[3 * j + k];
    T.m[3 * i + k] = c * a - s * b; T.m[3 * j + k] = s * a + c * b;
  }
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    float a = T.m[3 * k + i], b = T.m[3 * k + j];
    T.m[3 * k + i] = c * a - s * b; T.m[3 * k + j] = s * a + c * b;
  }
  return T;
}
template <int AX> GO2_HD M3 rot_sim_inv(float c, float s, const M3& M) { return rot_sim_fwd<AX>(c, -s, M); }  // E M E^T
template <int AX> GO2_HD M3 rot_left_fwd(float c, float s, const M3& M) {  // Rpc M
  M3 T = M;
  constexpr int i = AxPair<AX>::i, j = AxPair<AX>::j;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    float a = T.m[3 * i + k], b = T.m[3 * j + k];
    T.m[3 * i + k] = c * a - s * b; T.m[3 * j + k] = s * a + c * b;
  }
  return T;
}
GO2_HD M3 quat_to_mat(float x, float y, float z, float w) {
  M3 R;
  R.m[0] = 1 - 2 * (y * y + z * z); R.m[1] = 2 * (x * y - z * w); R.m[2] = 2 * (x * z + y * w);
  R.m[3] = 2 * (x * y + z * w); R.m[4] = 1 - 2 * (x * x + z * z); R.m[5] = 2 * (y * z - x * w);
  R.m[6] = 2 * (x * z - y * w); R.m[7] = 2 * (y * z + x * w); R.m[8] = 1 - 2 * (x * x + y * y);
  return R;
}
GO2_HD V3 quat_rotate_inverse(const float* q, V3 v) {  // isaacgym.torch_utils.quat_rotate_inverse, xyzw
  float w = q[3];
  V3 qv = mk(q[0], q[1], q[2]);
  V3 a = (2 * w * w - 1) * v, b = (2 * w) * cross(qv, v), c = (2 * dot(qv, v)) * qv;
  return mk(a.x - b.x + c.x, a.y - b.y + c.y, a.z - b.z + c.z);
}

// 6x6 symmetric operators in block form [A B; B^T C] (A, C symmetric but stored full)
struct Sym6 { M3 A, B, C; };
struct V6 { V3 a, l; };  // angular / linear parts (motion: w, v ; force: n, f)
GO2_HD V6 mul(const Sym6& I, const V6& v) { V6 r; r.a = mul(I.A, v.a) + mul(I.B, v.l); r.l = mulT(I.B, v.a) + mul(I.C, v.l); return r; }
GO2_HD void ld6(const float* p, V6& v) { v.a = ld3(p); v.l = ld3(p + 3); }
GO2_HD void st6(float* p, const V6& v) { st3(p, v.a); st3(p + 3, v.l); }
GO2_HD float comp(const V3& v, int k) { return k == 0 ? v.x : (k == 1 ? v.y : v.z); }
GO2_HD void addcomp(V3& v, int k, float d) { if (k == 0) v.x += d; else if (k == 1) v.y += d; else v.z += d; }
GO2_HD Sym6 rigid_inertia(const float* rec) {  // mass, com, Ixx Iyy Izz Ixy Ixz Iyz about COM -> about the link origin
  float m = rec[0];
  V3 c = mk(rec[1], rec[2], rec[3]);
  Sym6 I;
  float cc = dot(c, c);
  I.A.m[0] = rec[4] + m * (cc - c.x * c.x); I.A.m[1] = rec[7] - m * c.x * c.y; I.A.m[2] = rec[8] - m * c.x * c.z;
  I.A.m[3] = I.A.m[1]; I.A.m[4] = rec[5] + m * (cc - c.y * c.y); I.A.m[5] = rec[9] - m * c.y * c.z;
  I.A.m[6] = I.A.m[2]; I.A.m[7] = I.A.m[5]; I.A.m[8] = rec[6] + m * (cc - c.z * c.z);
  I.B.m[0] = 0; I.B.m[1] = -m * c.z; I.B.m[2] = m * c.y;
  I.B.m[3] = m * c.z; I.B.m[4] = 0; I.B.m[5] = -m * c.x;
  I.B.m[6] = -m * c.y; I.B.m[7] = m * c.x; I.B.m[8] = 0;
  for (int i = 0; i < 9; ++i) I.C.m[i] = 0;
  I.C.m[0] = I.C.m[4] = I.C.m[8] = m;
  return I;
}

// inverse of a symmetric positive definite 3x3 (adjugate / determinant; symmetric by construction)
GO2_HD M3 sym_inverse(const M3& M) {
  const float a = M.m[0], b = M.m[1], c = M.m[2], d = M.m[4], e = M.m[5], f = M.m[8];
  const float c00 = d * f - e * e, c01 = c * e - b * f, c02 = b * e - c * d;
  const float id = 1.0f / (a * c00 + b * c01 + c * c02);
  M3 R;
  R.m[0] = c00 * id; R.m[1] = c01 * id; R.m[2] = c02 * id;
  R.m[3] = R.m[1]; R.m[4] = (a * f - c * c) * id; R.m[5] = (b * c - a * e) * id;
  R.m[6] = R.m[2]; R.m[7] = R.m[5]; R.m[8] = (a * d - b * b) * id;
  return R;
}
// (I^A_0)^-1 of the floating base by 3x3 blocks: I = [A B; B^T C]  ->  [P Q; Q^T R] with the Schur complement of the linear block,
//   P = (A - B C^-1 B^T)^-1,  Q = -P B C^-1,  R = C^-1 - (B C^-1)^T Q.
// Two closed-form 3x3 inverses and four 3x3 products: a short dependency chain with wide instruction-level parallelism (the 6x6
// Cholesky + 12 triangular solves it replaces was a ~4.7 k-cycle serial chain of divisions and square roots per substep).
GO2_HD void base_inverse(const Sym6& I, M3& P, M3& Q, M3& R) {
  const M3 Ci = sym_inverse(I.C);
  const M3 T = mul(I.B, Ci);
  M3 S = sub(I.A, mul(T, transpose(I.B)));
  S.m[3] = S.m[1]; S.m[6] = S.m[2]; S.m[7] = S.m[5];      // exact symmetry (upper triangle wins)
  P = sym_inverse(S);
  const M3 PT = mul(P, T);
#pragma unroll
  for (int i = 0; i < 9; ++i) Q.m[i] = -PT.m[i];
  R = sub(Ci, mul(transpose(T), Q));
}

// ------------------------------------------------------------------------------------------------ per-warp scratch
template <int PAD>
struct WarpSmemT {
  float inertia[GO2_NUM_DYN * GO2_INERTIA_STRIDE];
  float Rw[GO2_NUM_DYN][9];
  float pw[GO2_NUM_DYN][3];
  float vs[GO2_NUM_DYN][6];    // body spatial velocity at the start of the substep (body coords)
  float v[GO2_NUM_DYN][6];     // unconstrained update vm (semi-implicit Euler, before impulses)
  float dv[GO2_NUM_DYN][6];    // response to the accumulated impulses
  float Lam[GO2_NUM_DYN][27];  // mobility blocks P, Q, R of bodies 1..12 (index 0 unused, base uses Lam0)
  float Lam0[36];              // (I^A_0)^-1, full 6x6
  float legIA[4][27], legpA[4][6], legp[4][6];
  union {                      // the contact phases of the substeps / the post-physics phases
    float fcol[GO2_NUM_COL][6];    // spatial impulse of each collider on its body (body coords)
    float heights[GO2_NUM_HEIGHT]; // height scan (written after the last substep)
  };
  float pcol[GO2_NUM_COL][3];  // world impulse of each collider
  float tgt[12][2];            // joint-limit target velocities (lower, upper row)
  float Dje[12];               // limit-row step limit_relax / (M^-1)_jj (limit_relax > 0)
  int bad;                     // state guard: this env's state went non-finite in this step (sanitised, resets)
  int stop_heading;            // heading commands: the yaw command no longer follows the heading target (legged_robot.py:412,431,548,582)
  float hrng[2];               // this env's heading range
  float mu_env, rest_env;      // contact friction / restitution of this env (combined with the terrain's)
  float a0[6];
  float q[12], qd[12], tau[12], cs[12][2], qdm[12], dqd[12], tauimp[12], Dj[12];
  float act[12], lact[12], llact[12], lqd[12], tq[12];
  float root[13];
  float cf[GO2_NUM_REPORT][3];
  float feet[4][6];
  float obsrow[76];            // proprioceptive columns of the privileged observation (go2_env.py:36-47), before clipping
  float part[32];
  float jterm[7][12];
  float fterm[4], coll[8];
  float blv[3], bav[3], pg[3];
  float cmd[4], cmd_rng[6];
  float termv[GO2_NUM_REW];
  float base_height, rew;
  float resamp_step, acc_xy[2], max_move;
  float env_origin[3];
  int active[GO2_NUM_COL];
  int ep_len, reset, tout, last_lim, level, ttype, tid, delay_start;
  int pad_[PAD];  // PAD = 6: stride = 1 mod 32 words, consecutive envs start one bank apart (the packed maps read 8 envs' scratch from one warp);
                  // PAD = 7: stride = 2 mod 32 words (half-warp maps: the two envs of a warp are 8 slots = 16 banks apart)
};
typedef WarpSmemT<6> WarpSmem;
static_assert((sizeof(WarpSmemT<6>) / 4) % 32 == 1, "WarpSmemT<6> stride must be 1 mod 32 words");
static_assert((sizeof(WarpSmemT<7>) / 4) % 32 == 2, "WarpSmemT<7> stride must be 2 mod 32 words");
// compile-time description of a kernel's thread map: virtual lanes per WIDE thread, role pruning, shared-memory stride
template <int NV_, int ROLE_, int PAD_> struct StepT { static constexpr int NV = NV_, ROLE = ROLE_; typedef WarpSmemT<PAD_> Smem; };

#define GO2_HALF_ENVS 14   /* envs per barrier group of the half-warp map */
struct Lane {
  // joint lanes (0..11)
  float kp, kd, mzo, mstr;
  float ddp, eff, qlo, qhi, vlim;   // default_dof_pos, effort / position / velocity limits of the joint (read once per step)
  // leg lanes (0..3): per link i of the leg
  float c[3], s[3], Dinv[3], u[3], uI[3];
  float U[3][6], cb[3][6];
  float pA[3][6];               // bias forces of pass 1, consumed by pass 2
  float lam_lo[3], lam_hi[3];   // accumulated joint-limit impulses of the leg's joints
  // collider items (one per virtual lane)
  struct Col { float n[3], Winv[6], vt, r[3], gsplit, gap; int body, act; } col[2];
  // thread map (init_roles): own env slot / lane, leg item, base-column item, first env id of the group, slots of the group
  int own, w, lane, nv;
  int leg, wl;
  int e0, w0, nw;
  // phase barrier: 0 = warp-level (__syncwarp); otherwise the number of threads of the CTA-wide named barrier every phase ends in
  // (required by the packed map; with the warp-per-env map it only keeps the warps of a CTA on the same stretch of code)
  int nsync;
  int nmid;      // same, at three more points inside a substep (GO2_MID_SYNC)
  int ncoarse;   // same, but only at the few GO2_COARSE_SYNC points (substep boundaries): loose re-alignment of the CTA's warps
};

#if defined(__CUDACC__)
#if defined(GO2_PHASE_TIMING)   // tuning build (tools/phase_timing.py): CTA GO2_PHASE_TIMING records clock64() after every phase barrier
__device__ long long go2_ph_clock[512];
__device__ int go2_ph_count;
#endif
__device__ __forceinline__ void go2_phase_sync(const Lane& L) {
  if (L.nsync) asm volatile("bar.sync 1, %0;" ::"r"(L.nsync) : "memory");
  else __syncwarp();
#if defined(GO2_PHASE_TIMING)
  if (threadIdx.x == 0 && blockIdx.x == GO2_PHASE_TIMING) { const int k = go2_ph_count; if (k < 512) go2_ph_clock[k] = clock64(); go2_ph_count = k + 1; }
#endif
}
// role-specialised kernels: the leg warps signal "kinematics done" without waiting (bar.arrive), the WIDE warps wait for it (bar.sync on the same
// named barrier); every thread of the leg warps arrives, also the ones without an item
#define GO2_KIN_ARRIVE() do { if (T::ROLE == 2) { __threadfence_block(); asm volatile("bar.arrive 4, %0;" ::"r"(L.nsync) : "memory"); } } while (0)
#define GO2_KIN_WAIT() do { asm volatile("bar.sync 4, %0;" ::"r"(L.nsync) : "memory"); } while (0)
#if defined(GO2_PHASE_TIMING)
#define GO2_TICK() do { if (threadIdx.x == 0 && blockIdx.x == GO2_PHASE_TIMING) { const int k = go2_ph_count; if (k < 512) go2_ph_clock[k] = clock64(); go2_ph_count = k + 1; } } while (0)
#else
#define GO2_TICK() do { } while (0)
#endif
#define GO2_COARSE_SYNC() do { if (L.ncoarse) asm volatile("bar.sync 2, %0;" ::"r"(L.ncoarse) : "memory"); } while (0)
#define GO2_MID_SYNC() do { if (L.nmid) asm volatile("bar.sync 3, %0;" ::"r"(L.nmid) : "memory"); } while (0)
#else
#define GO2_COARSE_SYNC() do { } while (0)
#define GO2_MID_SYNC() do { } while (0)
#define GO2_KIN_ARRIVE() do { } while (0)
#define GO2_KIN_WAIT() do { } while (0)
#define GO2_TICK() do { } while (0)
#endif

// Thread map of thread `tid` of a group of `nwarps` warps whose first env is e0 and which holds n_local (>= 1) envs.
//   packed == 0: warp w owns env e0 + w; its lanes 0..3 are that env's LEGS items.
//   packed == 1: 8 warps, up to 8 envs; warp w still owns env e0 + w for the WIDE role, but the LEGS items of all envs sit in
//                warp 0 (lane = 4 * slot + leg).
//                the LEGS items of all envs sit in ONE warp `leg_warp` (lane = 4 * slot + leg).  Which warp is free: the hardware maps warp w of a
//                CTA to scheduler w % 4, so co-resident CTAs that all used warp 0 would queue their serial leg streams on scheduler 0
//                (measured in round 1: scheduler 0 saturated, the other three a third busy); the kernels rotate it per SM.
GO2_HD void init_roles(Lane& L, int tid, int packed, int e0, int n_local, int nwarps, int leg_warp = 0) {
  const int warp = tid >> 5, lane = tid & 31;
  L.e0 = e0; L.w = warp; L.lane = lane; L.nv = 1;
  L.own = warp < n_local;
  L.ncoarse = 0; L.nmid = 0;
  if (!packed) {
    L.leg = (L.own && lane < 4) ? lane : -1; L.wl = warp;
    L.w0 = warp; L.nw = 1;
    L.nsync = 0;
  } else if (packed == 1) {
    L.leg = (warp == leg_warp && (lane >> 2) < n_local) ? (lane & 3) : -1; L.wl = lane >> 2;
    L.w0 = 0; L.nw = n_local;
    L.nsync = 32 * nwarps;
  } else {
    // packed == 2, the HALF-WARP map ("H14"): a group of 14 envs = 7 WIDE warps (16 threads per env, two virtual lanes each) + 2 dedicated
    // LEGS warps (56 (env, leg) items).  Warp w < 6 serves slots w and w + 8, warp 6 slots 6 and 7: with a scratch stride of 2 mod 32 words the
    // two halves of warps 0..5 sit 16 banks apart.  28 envs are resident per SM (two groups): 4096 envs are ONE wave on 148 SMs.
    const int nwide = GO2_HALF_ENVS / 2;
    if (warp < nwide) {
      const int h = lane >> 4;
      L.w = warp < nwide - 1 ? warp + 8 * h : nwide - 1 + h;
      L.lane = lane & 15; L.nv = 2;
      L.own = L.w < n_local;
      L.leg = -1; L.wl = 0;
    } else {
      const int it = (warp - nwide) * 32 + lane;
      L.w = 0; L.own = 0;
      L.wl = it >> 2;
      L.leg = (L.wl < n_local && L.wl < GO2_HALF_ENVS) ? (it & 3) : -1;
      if (L.leg < 0) L.wl = 0;
    }
    L.w0 = 0; L.nw = n_local;
    L.nsync = 32 * nwarps;
  }
}
// does any env of the thread's group reset this step?  (uniform over the threads that share phase barriers)
template <class SMT>
GO2_HD bool group_any_reset(const Lane& L, const SMT* SM) {
  bool any = false;
  for (int k = 0; k < L.nw; ++k) any = any || (SM[L.w0 + k].reset != 0);
  return any;
}

#if defined(__CUDACC__)
template <class T> __device__ __forceinline__ bool own_warp_any_reset(const Lane& L, const typename T::Smem* SM) {
  if (T::NV == 1) return SM[L.w].reset != 0;
  // the two slots of the thread's warp: its own and its partner's (slot pairs of init_roles, packed == 2)
  const int nwide = GO2_HALF_ENVS / 2, warp = threadIdx.x >> 5;
  const int s0 = warp < nwide - 1 ? warp : nwide - 1, s1 = warp < nwide - 1 ? warp + 8 : nwide;
  return (s0 < L.nw && SM[s0].reset != 0) || (s1 < L.nw && SM[s1].reset != 0);
}
#endif

struct StepCtx {
  const Go2EnvConfig* cfg; const Go2Model* mdl; const Go2EnvBuffers* buf; const Go2StepParams* sp;
  const float* actions_in;  // [N,12] or nullptr (reset_all / substeps entry points)
  const Go2EnvConfig* cs;   // where the SCALAR fields of the config are read from: the H14 kernel passes a by-value copy (kernel parameter space = constant
                            // bank operands, no load latency); the lane-indexed tables (kp, default_dof_pos, height_points ...) stay behind `cfg` (L1)
};

// joint axis of link index i within a leg: hip = x, thigh = calf = y (asserted on the host at create time)
template <int I> struct LinkAxis { static constexpr int ax = (I == 0) ? 0 : 1; };

#if !defined(__CUDACC__)
static inline int min(int a, int b) { return a < b ? a : b; }
static inline int max(int a, int b) { return a > b ? a : b; }
#endif

// bilinear heightfield query (plane: every sample reads as 0, h = 0).  Branch-free: the loads are predicated, so the queries of a thread's two
// virtual lanes overlap
#if defined(__CUDACC__)
__device__ __forceinline__ float ldg_h_if(bool p, const int16_t* a) {
  int v = 0;
  asm volatile("{\n .reg .pred q;\n setp.ne.b32 q, %2, 0;\n @q ld.global.nc.s16 %0, [%1];\n}" : "+r"(v) : "l"(a), "r"((int)p));
  return (float)v;
}
#else
static inline float ldg_h_if(bool p, const int16_t* a) { return p ? (float)*a : 0.0f; }
#endif
GO2_HD void terrain_query(const Go2EnvConfig* C, const int16_t* hs, float x, float y, float& h, float& dhx, float& dhy) {
  const bool hf = C->mesh_type != 0;
  float gx = (x + C->border) / C->hscale, gy = (y + C->border) / C->hscale;
  int ix = (int)floorf(gx), iy = (int)floorf(gy);
  ix = min(max(ix, 0), C->hf_rows - 2);
  iy = min(max(iy, 0), C->hf_cols - 2);
  float fx = fminf(fmaxf(gx - (float)ix, 0.0f), 1.0f), fy = fminf(fmaxf(gy - (float)iy, 0.0f), 1.0f);
  const int16_t* p = hs + (size_t)ix * C->hf_cols + iy;
  float h00 = ldg_h_if(hf, p), h01 = ldg_h_if(hf, p + 1), h10 = ldg_h_if(hf, p + C->hf_cols), h11 = ldg_h_if(hf, p + C->hf_cols + 1);
  float vs = C->vscale, k = vs / C->hscale;
  h = vs * ((1 - fx) * (1 - fy) * h00 + fx * (1 - fy) * h10 + (1 - fx) * fy * h01 + fx * fy * h11);
  dhx = k * ((1 - fy) * (h10 - h00) + fy * (h11 - h01));
  dhy = k * ((1 - fx) * (h01 - h00) + fx * (h11 - h10));
  if (!hf) { h = 0; dhx = 0; dhy = 0; }
}


// ================================================================================================ leg-lane routines
// Pass 1 + 2 of the articulated-body algorithm for the three links of leg l (hip, thigh, calf), leaves to root.
// Leaves U, Dinv, u, c (bias accel), (c,s) in the Lane; writes world transforms, start velocities and the leg's
// contribution to the base's articulated inertia / bias force into shared memory.
template <int I, class SMT>
GO2_HD void leg_pass1(int l, Lane& L, SMT& S, const Go2Model* M, V6& vpar, M3& Rwp, V3& pwp, V6 (&vl)[3], V6 (&pA)[3]) {
  constexpr int AX = LinkAxis<I>::ax;
  const int j = 3 * l + I, b = j + 1;
  float c = S.cs[j][0], s = S.cs[j][1];
  L.c[I] = c; L.s[I] = s;
  V3 r = ld3(M->joint_origin[j]);
  // world transform
  V3 pw = pwp + mul(Rwp, r);
  M3 Rw = Rwp;  // Rw_child = Rwp * Rpc : rotate columns (i, j) of Rwp
  {
    constexpr int ci = AxPair<AX>::i, cj = AxPair<AX>::j;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      float a = Rwp.m[3 * k + ci], bb = Rwp.m[3 * k + cj];
      Rw.m[3 * k + ci] = c * a + s * bb; Rw.m[3 * k + cj] = -s * a + c * bb;
    }
  }
  stm(S.Rw[b], Rw); st3(S.pw[b], pw);
  // velocity: v = X vpar + S qd
  V6 v;
  v.a = rot_inv<AX>(c, s, vpar.a);
  v.l = rot_inv<AX>(c, s, vpar.l + cross(vpar.a, r));
  float qd = S.qd[j];
  addcomp(v.a, AX, qd);
  st6(S.vs[b], v);
  // bias acceleration c = v x (S qd)
  V3 e = mk(AX == 0 ? qd : 0.0f, AX == 1 ? qd : 0.0f, AX == 2 ? qd : 0.0f);
  V3 ca = cross(v.a, e), cl = cross(v.l, e);
  L.cb[I][0] = ca.x; L.cb[I][1] = ca.y; L.cb[I][2] = ca.z; L.cb[I][3] = cl.x; L.cb[I][4] = cl.y; L.cb[I][5] = cl.z;
  // bias force pA = v x* (I v)
  Sym6 Ib = rigid_inertia(S.inertia + b * GO2_INERTIA_STRIDE);
  V6 h = mul(Ib, v);
  pA[I].a = cross(v.a, h.a) + cross(v.l, h.l);
  pA[I].l = cross(v.a, h.l);
  vl[I] = v;
  vpar = v; Rwp = Rw; pwp = pw;
}

// one backward step: consumes the articulated inertia IA / bias pA of link I, returns their contribution to the parent
template <int I, class SMT>
GO2_HD void leg_pass2(int l, Lane& L, SMT& S, const Go2Model* M, Sym6& IA, V6& pA, Sym6& IAout, V6& pAout) {
  constexpr int AX = LinkAxis<I>::ax;
  const int j = 3 * l + I;
  float c = L.c[I], s = L.s[I];
  V3 r = ld3(M->joint_origin[j]);
  // U = IA[:, k] : (A[:,k], B[k,:])
  V3 Ua = mk(IA.A.m[AX], IA.A.m[3 + AX], IA.A.m[6 + AX]);
  V3 Ul = mk(IA.B.m[3 * AX], IA.B.m[3 * AX + 1], IA.B.m[3 * AX + 2]);
  float D = comp(Ua, AX), Dinv = 1.0f / D;
  float u = S.tau[j] - comp(pA.a, AX);
  L.Dinv[I] = Dinv; L.u[I] = u;
  S.Dj[j] = D;
  L.U[I][0] = Ua.x; L.U[I][1] = Ua.y; L.U[I][2] = Ua.z; L.U[I][3] = Ul.x; L.U[I][4] = Ul.y; L.U[I][5] = Ul.z;
  // Ia = IA - U U^T / D
  float ua[3] = {Ua.x, Ua.y, Ua.z}, ul[3] = {Ul.x, Ul.y, Ul.z};
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int b = 0; b < 3; ++b) {
      IA.A.m[3 * a + b] -= ua[a] * ua[b] * Dinv;
      IA.B.m[3 * a + b] -= ua[a] * ul[b] * Dinv;
      IA.C.m[3 * a + b] -= ul[a] * ul[b] * Dinv;
    }
  // pa = pA + Ia c + U u / D
  V6 cb; cb.a = mk(L.cb[I][0], L.cb[I][1], L.cb[I][2]); cb.l = mk(L.cb[I][3], L.cb[I][4], L.cb[I][5]);
  V6 Ic = mul(IA, cb);
  float ud = u * Dinv;
  V6 pa; pa.a = pA.a + Ic.a + ud * Ua; pa.l = pA.l + Ic.l + ud * Ul;
  // to the parent frame: rotate blocks by Rpc, then shift by r
  M3 Ab = rot_sim_fwd<AX>(c, s, IA.A), Bb = rot_sim_fwd<AX>(c, s, IA.B), Cb = rot_sim_fwd<AX>(c, s, IA.C);
  // A_p = Ab + rx Bb^T - Bb rx - rx Cb rx ; B_p = Bb + rx Cb ; C_p = Cb
  M3 rxC = skew_mul(r, Cb);
  IAout.C = Cb;
  IAout.B = add(Bb, rxC);
  M3 rxBt = skew_mul(r, transpose(Bb));
  M3 Brx = mul_skew(Bb, r);
  M3 rxCrx = mul_skew(rxC, r);
  IAout.A = sub(sub(add(Ab, rxBt), Brx), rxCrx);
  V3 fb = rot_fwd<AX>(c, s, pa.l), nb = rot_fwd<AX>(c, s, pa.a);
  pAout.l = fb; pAout.a = nb + cross(r, fb);
}

// impulse response, inward: p (force-space, body coords) of link I -> contribution to the parent; stores uI
template <int I, class SMT>
GO2_HD void leg_imp_in(int l, Lane& L, const SMT& S, const Go2Model* M, const V6& p, V6& pout) {
  constexpr int AX = LinkAxis<I>::ax;
  const int j = 3 * l + I;
  float uI = S.tauimp[j] - comp(p.a, AX);
  L.uI[I] = uI;
  float ud = uI * L.Dinv[I];
  V6 pa;
  pa.a = p.a + ud * mk(L.U[I][0], L.U[I][1], L.U[I][2]);
  pa.l = p.l + ud * mk(L.U[I][3], L.U[I][4], L.U[I][5]);
  V3 r = ld3(M->joint_origin[j]);
  V3 fb = rot_fwd<AX>(L.c[I], L.s[I], pa.l), nb = rot_fwd<AX>(L.c[I], L.s[I], pa.a);
  pout.l = fb; pout.a = nb + cross(r, fb);
}
// outward: parent's dv -> this link's dv and joint velocity change
template <int I, class SMT>
GO2_HD void leg_imp_out(int l, Lane& L, SMT& S, const Go2Model* M, V6& dvpar) {
  constexpr int AX = LinkAxis<I>::ax;
  const int j = 3 * l + I, b = j + 1;
  V3 r = ld3(M->joint_origin[j]);
  V6 dp;
  dp.a = rot_inv<AX>(L.c[I], L.s[I], dvpar.a);
  dp.l = rot_inv<AX>(L.c[I], L.s[I], dvpar.l + cross(dvpar.a, r));
  float Ud = L.U[I][0] * dp.a.x + L.U[I][1] * dp.a.y + L.U[I][2] * dp.a.z + L.U[I][3] * dp.l.x + L.U[I][4] * dp.l.y + L.U[I][5] * dp.l.z;
  float dq = (L.uI[I] - Ud) * L.Dinv[I];
  addcomp(dp.a, AX, dq);
  S.dqd[j] = dq;
  st6(S.dv[b], dp);
  dvpar = dp;
}
// pass 3 + unconstrained velocity + mobility, outward for link I
template <int I, class SMT>
GO2_HD void leg_pass3(int l, Lane& L, SMT& S, const Go2Model* M, float dt, float limit_relax, V6& apar, V6& vmpar, M3& Pp, M3& Qp, M3& Rp) {
  constexpr int AX = LinkAxis<I>::ax;
  const int j = 3 * l + I, b = j + 1;
  float c = L.c[I], s = L.s[I];
  V3 r = ld3(M->joint_origin[j]);
  V3 Ua = mk(L.U[I][0], L.U[I][1], L.U[I][2]), Ul = mk(L.U[I][3], L.U[I][4], L.U[I][5]);
  float Dinv = L.Dinv[I];
  // acceleration
  V6 ap;
  ap.a = rot_inv<AX>(c, s, apar.a) + mk(L.cb[I][0], L.cb[I][1], L.cb[I][2]);
  ap.l = rot_inv<AX>(c, s, apar.l + cross(apar.a, r)) + mk(L.cb[I][3], L.cb[I][4], L.cb[I][5]);
  float qdd = (L.u[I] - (dot(Ua, ap.a) + dot(Ul, ap.l))) * Dinv;
  addcomp(ap.a, AX, qdd);
  apar = ap;
  float qdm = S.qd[j] + dt * qdd;
  S.qdm[j] = qdm;
  // unconstrained velocity of this link, consistent with the configuration
  V6 vm;
  vm.a = rot_inv<AX>(c, s, vmpar.a);
  vm.l = rot_inv<AX>(c, s, vmpar.l + cross(vmpar.a, r));
  addcomp(vm.a, AX, qdm);
  st6(S.v[b], vm);
  vmpar = vm;
  // mobility: Lam = L^T (X Lam_p X^T) L + S S^T / D.  Shift by r, rotate by E, then the rank-one row/column update.
  // shift: P' = P ; Q' = P rx + Q ; R' = R - rx Q + Q^T rx - rx P rx
  M3 Prx = mul_skew(Pp, r);
  M3 Qs = add(Prx, Qp);
  M3 rxQ = skew_mul(r, Qp);
  M3 Qtrx = mul_skew(transpose(Qp), r);
  M3 rxPrx = skew_mul(r, Prx);
  M3 Rs = sub(add(sub(Rp, rxQ), Qtrx), rxPrx);
  M3 P = rot_sim_inv<AX>(c, s, Pp), Q = rot_sim_inv<AX>(c, s, Qs), R = rot_sim_inv<AX>(c, s, Rs);
  // y = Mmat U ; alpha = U^T y
  V3 ya = mul(P, Ua) + mul(Q, Ul), yl = mulT(Q, Ua) + mul(R, Ul);
  float alpha = dot(Ua, ya) + dot(Ul, yl);
  float yav[3] = {ya.x, ya.y, ya.z}, ylv[3] = {yl.x, yl.y, yl.z};
#pragma unroll
  for (int t = 0; t < 3; ++t) {
    P.m[3 * AX + t] -= yav[t] * Dinv;   // row k of P
    P.m[3 * t + AX] -= yav[t] * Dinv;   // column k of P
    Q.m[3 * AX + t] -= ylv[t] * Dinv;   // row k of Q (= column k of Q^T)
  }
  P.m[3 * AX + AX] += alpha * Dinv * Dinv + Dinv;
  if (limit_relax > 0.0f) S.Dje[j] = limit_relax / P.m[3 * AX + AX];   // P[k][k] = (M^-1)_jj: the joint's own response with every other joint free
  stm(S.Lam[b], P); stm(S.Lam[b] + 9, Q); stm(S.Lam[b] + 18, R);
  Pp = P; Qp = Q; Rp = R;
}

// ================================================================================================ one physics substep
// `lane`/`L` come from the enclosing GO2_LANES_BEGIN; S is this warp's scratch.  last = last substep (report forces).
#if defined(__CUDACC__)
#define GO2_LANE_ARGS Lane& L
#define GO2_LANE_PASS L
/* the env's own warp decides alone: reset_phases holds WIDE phases only (half-warp maps: either env of the warp, read through the warp's first thread's view) */
#define GO2_ANY_RESET(SM) (T::ROLE != 2 && own_warp_any_reset<T>(L, SM))
#else
#define GO2_LANE_ARGS Lane* lanes, int NT
#define GO2_LANE_PASS lanes, NT
#define GO2_ANY_RESET(SM) group_any_reset(lanes[0], SM)   /* the emulated group walks reset_phases when any of its envs resets */
#endif

// ---- S6a: collider lanes: narrow phase against the terrain (needs the world transforms of ABA pass 1 only).  Branch-free, so that the two
// virtual lanes of a half-warp thread overlap their heightfield loads.  In the role-specialised kernels the WIDE warps run it WHILE the leg warps
// are in ABA passes 2-3 (GO2_KIN_ARRIVE / GO2_KIN_WAIT below).
template <class T>
GO2_HD void narrow_phase(GO2_LANE_ARGS, typename T::Smem* SM, const StepCtx& X) {
  const Go2EnvConfig* C = X.cs;
  const Go2Model* M = X.mdl;
  GO2_WIDE {
    {
      Lane::Col& K = L.col[vh];
      const int ci = lane;
      const int b = M->col_dyn[ci];
      K.body = b;
      V3 r = ld3(M->col_pos[ci]);
      K.r[0] = r.x; K.r[1] = r.y; K.r[2] = r.z;
      M3 Rw = ldm(S.Rw[b]);
      V3 cw = ld3(S.pw[b]) + mul(Rw, r);
      float h, dhx, dhy;
      terrain_query(C, X.buf->height_samples, cw.x, cw.y, h, dhx, dhy);
      float inv = 1.0f / sqrtf(dhx * dhx + dhy * dhy + 1.0f);
      V3 n = mk(-dhx * inv, -dhy * inv, inv);
      float gap = (cw.z - h) * n.z - M->col_radius[ci];
      int act = gap < C->contact_offset;
      K.act = act; S.active[ci] = act; K.gap = gap;
      K.n[0] = n.x; K.n[1] = n.y; K.n[2] = n.z;
      for (int k = 0; k < 6; ++k) S.fcol[ci][k] = 0;
      S.pcol[ci][0] = S.pcol[ci][1] = S.pcol[ci][2] = 0;
    }
  } GO2_SYNC_WARP();
}

template <class T>
GO2_HD void physics_substep(GO2_LANE_ARGS, typename T::Smem* SM, const StepCtx& X, bool last) {
  const Go2EnvConfig* C = X.cs;
  const Go2EnvConfig* CT = X.cfg; (void)CT;
  const Go2Model* M = X.mdl;
  const float dt = C->sim_dt;
  // ---- S1: joint lanes: sin/cos, limit targets; lane 12: base kinematics
  GO2_WIDE {
    if (lane < GO2_NUM_DOF) {
      float q = S.q[lane];
      float sn, cn;
#if defined(__CUDACC__)
      sincosf(q, &sn, &cn);
#else
      sn = sinf(q); cn = cosf(q);
#endif
      S.cs[lane][0] = cn; S.cs[lane][1] = sn;
      float glo = q - L.qlo, ghi = L.qhi - q;
      S.tgt[lane][0] = (glo >= 0) ? -glo / dt : -glo * C->limit_erp / dt;
      S.tgt[lane][1] = (ghi >= 0) ? ghi / dt : ghi * C->limit_erp / dt;
      S.dqd[lane] = 0; S.tauimp[lane] = 0;
    }
    if (lane == 12) {
      M3 R0 = quat_to_mat(S.root[3], S.root[4], S.root[5], S.root[6]);
      stm(S.Rw[0], R0);
      st3(S.pw[0], ld3(S.root));
      V3 wb = mulT(R0, ld3(S.root + 10)), vb = mulT(R0, ld3(S.root + 7));
      st3(S.vs[0], wb); st3(S.vs[0] + 3, vb);
    }
    if (lane >= 13 && lane < 13 + GO2_NUM_DYN) { for (int k = 0; k < 6; ++k) S.dv[lane - 13][k] = 0; }
  } GO2_SYNC();
  // ---- S2: leg lanes: ABA pass 1 (kinematics, bias forces) ...
  GO2_LEGS {
    if (lane < 4) {
      V6 vpar; ld6(S.vs[0], vpar);
      M3 Rwp = ldm(S.Rw[0]); V3 pwp = ld3(S.pw[0]);
      V6 vl[3], pA[3];
      leg_pass1<0>(lane, L, S, M, vpar, Rwp, pwp, vl, pA);
      leg_pass1<1>(lane, L, S, M, vpar, Rwp, pwp, vl, pA);
      leg_pass1<2>(lane, L, S, M, vpar, Rwp, pwp, vl, pA);
      for (int i = 0; i < 3; ++i) st6(L.pA[i], pA[i]);
    }
  }
  GO2_KIN_ARRIVE();   // the world transforms are in shared memory: the WIDE warps of a role-specialised kernel start the narrow phase (every thread of the leg warps arrives)
  // ---- ... and pass 2 (articulated inertias, leaves to root)
  GO2_LEGS {
    if (lane < 4) {
      V6 pA[3];
      for (int i = 0; i < 3; ++i) ld6(L.pA[i], pA[i]);
      Sym6 IA = rigid_inertia(S.inertia + (3 * lane + 3) * GO2_INERTIA_STRIDE), Iout;
      V6 pout;
      leg_pass2<2>(lane, L, S, M, IA, pA[2], Iout, pout);
      IA = rigid_inertia(S.inertia + (3 * lane + 2) * GO2_INERTIA_STRIDE);
      IA.A = add(IA.A, Iout.A); IA.B = add(IA.B, Iout.B); IA.C = add(IA.C, Iout.C);
      pA[1].a = pA[1].a + pout.a; pA[1].l = pA[1].l + pout.l;
      leg_pass2<1>(lane, L, S, M, IA, pA[1], Iout, pout);
      IA = rigid_inertia(S.inertia + (3 * lane + 1) * GO2_INERTIA_STRIDE);
      IA.A = add(IA.A, Iout.A); IA.B = add(IA.B, Iout.B); IA.C = add(IA.C, Iout.C);
      pA[0].a = pA[0].a + pout.a; pA[0].l = pA[0].l + pout.l;
      leg_pass2<0>(lane, L, S, M, IA, pA[0], Iout, pout);
      stm(S.legIA[lane], Iout.A); stm(S.legIA[lane] + 9, Iout.B); stm(S.legIA[lane] + 18, Iout.C);
      st6(S.legpA[lane], pout);
    }
  }
  // ---- S3 + S5: leg lanes.  Every leg lane of an env assembles I^A_0 and inverts it (same instruction stream on the 4 lanes, no
  // cross-lane traffic), then runs pass 3 / unconstrained velocities / the mobility recursion for its own leg.  The only dependency on
  // the other legs is their legIA / legpA: a warp-level sync (the 4 leg lanes of an env always share a warp).
  GO2_SYNC_WARP();
  if (T::ROLE == 1) { GO2_KIN_WAIT(); narrow_phase<T>(GO2_LANE_PASS, SM, X); }
  GO2_LEGS {
    if (lane < 4) {
      Sym6 I0 = rigid_inertia(S.inertia);
      V6 v0; ld6(S.vs[0], v0);
      V6 h = mul(I0, v0);
      V6 p0; p0.a = cross(v0.a, h.a) + cross(v0.l, h.l); p0.l = cross(v0.a, h.l);
      for (int l = 3; l >= 0; --l) {  // same order as the oracle's leaves-to-root sweep
        I0.A = add(I0.A, ldm(S.legIA[l])); I0.B = add(I0.B, ldm(S.legIA[l] + 9)); I0.C = add(I0.C, ldm(S.legIA[l] + 18));
        V6 pl; ld6(S.legpA[l], pl);
        p0.a = p0.a + pl.a; p0.l = p0.l + pl.l;
      }
      M3 P, Q, R;
      base_inverse(I0, P, Q, R);
      V6 a0;   // a0 = -Lam0 p0
      {
        V3 ta = mul(P, p0.a) + mul(Q, p0.l), tl = mulT(Q, p0.a) + mul(R, p0.l);
        a0.a = mk(-ta.x, -ta.y, -ta.z); a0.l = mk(-tl.x, -tl.y, -tl.z);
      }
      if (lane == 0) {   // the base's mobility for the contact phases (colliders on body 0, impulse response of the base)
        for (int i = 0; i < 3; ++i)
          for (int j = 0; j < 3; ++j) {
            S.Lam0[6 * i + j] = P.m[3 * i + j]; S.Lam0[6 * i + j + 3] = Q.m[3 * i + j];
            S.Lam0[6 * (i + 3) + j] = Q.m[3 * j + i]; S.Lam0[6 * (i + 3) + j + 3] = R.m[3 * i + j];
          }
      }
      M3 R0 = ldm(S.Rw[0]);
      V3 gb = mulT(R0, mk(0.0f, 0.0f, C->gravity_z));
      V6 vm0;
      vm0.a = v0.a + dt * a0.a;
      vm0.l = v0.l + dt * (a0.l + gb + cross(v0.a, v0.l));  // components stay in the frame of the start of the step
      V6 ap = a0, vmp = vm0;
      leg_pass3<0>(lane, L, S, M, dt, C->limit_relax, ap, vmp, P, Q, R);
      leg_pass3<1>(lane, L, S, M, dt, C->limit_relax, ap, vmp, P, Q, R);
      leg_pass3<2>(lane, L, S, M, dt, C->limit_relax, ap, vmp, P, Q, R);
      if (lane == 0) st6(S.v[0], vm0);
      for (int i = 0; i < 3; ++i) { L.lam_lo[i] = 0; L.lam_hi[i] = 0; }
    }
  } GO2_SYNC();
  // ---- S6a (narrow phase) for the thread maps whose threads carry both roles; the role-specialised kernels ran it above, under the leg warps
  if (T::ROLE != 1) narrow_phase<T>(GO2_LANE_PASS, SM, X);
  // ---- S6b: collider lanes: per-contact 3x3 mobility and velocity target of the active colliders
  GO2_WIDE {
    {
      Lane::Col& K = L.col[vh];
      const int ci = lane, b = K.body;
      if (K.act) {
        V3 r = mk(K.r[0], K.r[1], K.r[2]), n = mk(K.n[0], K.n[1], K.n[2]);
        M3 Rw = ldm(S.Rw[b]);
        M3 P, Q, R;
        if (b == 0) {
          for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) { P.m[3 * i + j] = S.Lam0[6 * i + j]; Q.m[3 * i + j] = S.Lam0[6 * i + j + 3]; R.m[3 * i + j] = S.Lam0[6 * (i + 3) + j + 3]; }
        } else { P = ldm(S.Lam[b]); Q = ldm(S.Lam[b] + 9); R = ldm(S.Lam[b] + 18); }
        // W = Rw (R - rx Q + Q^T rx - rx P rx) Rw^T : mobility of the point, world frame
        M3 Prx = mul_skew(P, r);
        M3 Wb = sub(add(sub(R, skew_mul(r, Q)), mul_skew(transpose(Q), r)), skew_mul(r, Prx));
        M3 W = mul(mul(Rw, Wb), transpose(Rw));
        float a = W.m[0], bq = W.m[1], c2 = W.m[2], d = W.m[4], e2 = W.m[5], f = W.m[8];
        float c00 = d * f - e2 * e2, c01 = c2 * e2 - bq * f, c02 = bq * e2 - c2 * d;
        float id = 1.0f / (a * c00 + bq * c01 + c2 * c02);
        K.Winv[0] = c00 * id; K.Winv[1] = c01 * id; K.Winv[2] = c02 * id;
        K.Winv[3] = (a * f - c2 * c2) * id; K.Winv[4] = (bq * c2 - a * e2) * id; K.Winv[5] = (a * d - bq * bq) * id;
        // normal velocity target; restitution looks at the approach speed at the START of the step
        V6 vsb; ld6(S.vs[b], vsb);
        float vn0 = dot(mul(Rw, vsb.l + cross(vsb.a, r)), n);
        const float gap = K.gap;
        float vt = (gap >= 0) ? -gap / dt : fminf(fmaxf(-gap - C->penetration_slop, 0.0f) * C->erp / dt, C->max_depen_vel);
        if (vn0 < -C->bounce_threshold) vt = fmaxf(vt, -S.rest_env * vn0);
        K.vt = vt;
      }
      (void)ci;
    }
  } GO2_SYNC_WARP();
  // ---- S7: mass-splitting factor of the collider's group (base = colliders 0..7, leg l = 8+6l .. 13+6l)
  GO2_WIDE {
    {
      Lane::Col& K = L.col[vh];
      int g0 = lane < 8 ? 0 : 8 + 6 * ((lane - 8) / 6), gn = lane < 8 ? 8 : 6, cnt = 0;
      for (int k = 0; k < gn; ++k) cnt += S.active[g0 + k];
      K.gsplit = C->contact_relax / (float)cnt;      // block step of the contact rows; only read by active colliders: cnt >= 1
    }
  }
  // ---- Jacobi sweeps with exact propagation through the tree: [collider lanes: block-solve every contact] | CTA barrier |
  // [leg threads: limit rows, impulses inward, base response, outward] | CTA barrier
  GO2_MID_SYNC();
  for (int it = 0; it < C->solver_iters; ++it) {
    GO2_WIDE {
      Lane::Col& K = L.col[vh];
      if (K.act) {
        const int b = K.body;
        V3 r = mk(K.r[0], K.r[1], K.r[2]), n = mk(K.n[0], K.n[1], K.n[2]);
        M3 Rw = ldm(S.Rw[b]);
        V6 vb, db; ld6(S.v[b], vb); ld6(S.dv[b], db);
        vb.a = vb.a + db.a; vb.l = vb.l + db.l;
        V3 vp = mul(Rw, vb.l + cross(vb.a, r));
        V3 err = vp - K.vt * n;
        float is = K.gsplit;
        V3 we = mk(K.Winv[0] * err.x + K.Winv[1] * err.y + K.Winv[2] * err.z, K.Winv[1] * err.x + K.Winv[3] * err.y + K.Winv[4] * err.z,
                   K.Winv[2] * err.x + K.Winv[4] * err.y + K.Winv[5] * err.z);
        V3 pc = ld3(S.pcol[lane]) - is * we;   // the accumulated impulse lives in pcol (zeroed by the narrow phase)
        float pcn = dot(pc, n);
        float pn = fmaxf(0.0f, pcn);
        V3 pt = pc - pcn * n;
        float ptn = sqrtf(dot(pt, pt)), lim = S.mu_env * pn;
        if (ptn > lim) pt = (ptn > 0 ? lim / ptn : 0.0f) * pt;
        V3 p = pn * n + pt;
        V3 fl = mulT(Rw, p), fn = cross(r, fl);
        st3(S.fcol[lane], fn); st3(S.fcol[lane] + 3, fl);
        st3(S.pcol[lane], p);
      }
    } GO2_SYNC();
    GO2_LEGS {
      if (lane < 4) {
        for (int i = 0; i < 3; ++i) {  // joint-limit rows of the leg's joints (unilateral, velocity level)
          const int j = 3 * lane + i;
          float cur = S.qdm[j] + S.dqd[j], Dj = C->limit_relax > 0.0f ? S.Dje[j] : S.Dj[j];
          L.lam_lo[i] = fmaxf(0.0f, L.lam_lo[i] + (S.tgt[j][0] - cur) * Dj);
          L.lam_hi[i] = fminf(0.0f, L.lam_hi[i] + (S.tgt[j][1] - cur) * Dj);
          S.tauimp[j] = L.lam_lo[i] + L.lam_hi[i];
        }
        // inward: gather the leg's collider impulses (fixed order) and push them to the base
        const int c0 = 8 + 6 * lane;
        V6 p2, p1, p0, out;
        p2.a = mk(0, 0, 0); p2.l = mk(0, 0, 0);
        for (int k = 2; k < 6; ++k) { p2.a = p2.a - ld3(S.fcol[c0 + k]); p2.l = p2.l - ld3(S.fcol[c0 + k] + 3); }
        leg_imp_in<2>(lane, L, S, M, p2, out);
        p1.a = out.a - ld3(S.fcol[c0 + 1]); p1.l = out.l - ld3(S.fcol[c0 + 1] + 3);
        leg_imp_in<1>(lane, L, S, M, p1, out);
        p0.a = out.a - ld3(S.fcol[c0]); p0.l = out.l - ld3(S.fcol[c0] + 3);
        leg_imp_in<0>(lane, L, S, M, p0, out);
        st6(S.legp[lane], out);
      }
    } GO2_SYNC_WARP();
    // base: dv0 = -Lam0 p0, evaluated by every leg lane of the env (same stream, no extra phase), then outward along the leg
    GO2_LEGS {
      if (lane < 4) {
        float p0[6] = {0, 0, 0, 0, 0, 0};
        for (int c = 0; c < 8; ++c) for (int k = 0; k < 6; ++k) p0[k] -= S.fcol[c][k];
        for (int l = 3; l >= 0; --l) for (int k = 0; k < 6; ++k) p0[k] += S.legp[l][k];
        float d0[6];
        for (int i = 0; i < 6; ++i) {
          float acc = 0;
          for (int k = 0; k < 6; ++k) acc += S.Lam0[6 * i + k] * p0[k];
          d0[i] = -acc;
        }
        V6 dvp; dvp.a = mk(d0[0], d0[1], d0[2]); dvp.l = mk(d0[3], d0[4], d0[5]);
        if (lane == 0) st6(S.dv[0], dvp);
        leg_imp_out<0>(lane, L, S, M, dvp);
        leg_imp_out<1>(lane, L, S, M, dvp);
        leg_imp_out<2>(lane, L, S, M, dvp);
      }
    } GO2_SYNC();
  }
  // ---- S12: final velocities, joint velocity clamp, integration, contact force report
  GO2_MID_SYNC();
  GO2_WIDE {
    if (lane < GO2_NUM_DOF) {
      float x = S.qdm[lane] + S.dqd[lane], vl = L.vlim;
      x = fminf(fmaxf(x, -vl), vl);
      S.qd[lane] = x;
      S.q[lane] += dt * x;
    }
    if (lane == 12) {
      M3 R0 = ldm(S.Rw[0]);
      V6 v0; ld6(S.v[0], v0);
      V6 d0; ld6(S.dv[0], d0);
      V3 lw = mul(R0, v0.l + d0.l), aw = mul(R0, v0.a + d0.a);
      if (C->state_guard) {   // asset.max_linear_velocity / max_angular_velocity (legged_robot_config.py:131-132)
        const float nl = sqrtf(dot(lw, lw)), na = sqrtf(dot(aw, aw));
        if (nl > C->max_base_lin_vel) lw = (C->max_base_lin_vel / nl) * lw;
        if (na > C->max_base_ang_vel) aw = (C->max_base_ang_vel / na) * aw;
      }
      st3(S.root + 7, lw); st3(S.root + 10, aw);
      S.root[0] += dt * lw.x; S.root[1] += dt * lw.y; S.root[2] += dt * lw.z;
      float wn = sqrtf(dot(aw, aw)), th = wn * dt, hx, hy, hz, hw;  // q <- exp(aw dt / 2) * q (world-frame angular velocity)
      if (th > 1e-8f) {
        float sh = sinf(0.5f * th) / wn;
        hx = aw.x * sh; hy = aw.y * sh; hz = aw.z * sh; hw = cosf(0.5f * th);
      } else { hx = 0.5f * dt * aw.x; hy = 0.5f * dt * aw.y; hz = 0.5f * dt * aw.z; hw = 1.0f; }
      float x = S.root[3], y = S.root[4], z = S.root[5], w = S.root[6];
      float nx = hw * x + hx * w + hy * z - hz * y, ny = hw * y - hx * z + hy * w + hz * x;
      float nz = hw * z + hx * y - hy * x + hz * w, nw = hw * w - hx * x - hy * y - hz * z;
      float nn = 1.0f / sqrtf(nx * nx + ny * ny + nz * nz + nw * nw);
      S.root[3] = nx * nn; S.root[4] = ny * nn; S.root[5] = nz * nn; S.root[6] = nw * nn;
    }
    if (last && lane >= 13) {  // 19 reported bodies on lanes 13..31
      const int b = lane - 13;
      int c0, n;
      if (b == 0) { c0 = 0; n = 6; }
      else if (b <= 2) { c0 = 5 + b; n = 1; }
      else {
        int l = (b - 3) / 4, k = (b - 3) % 4;   // hip, thigh, calf, foot
        c0 = 8 + 6 * l + (k == 0 ? 0 : (k == 1 ? 1 : (k == 2 ? 2 : 5)));
        n = (k == 2) ? 3 : 1;
      }
      V3 f = mk(0, 0, 0);
      for (int k = 0; k < n; ++k) f = f + ld3(S.pcol[c0 + k]);
      float idt = 1.0f / dt;
      S.cf[b][0] = f.x * idt; S.cf[b][1] = f.y * idt; S.cf[b][2] = f.z * idt;
    }
  } GO2_SYNC_WARP();
}

// State guard (Go2EnvConfig.state_guard): an env whose state is non-finite after the substeps restarts from its initial pose at its origin and
// resets in this step, so that nothing non-finite reaches the observations / rewards the shared networks train on.  One WIDE phase, lane 0.
template <class T>
GO2_HD void state_guard(GO2_LANE_ARGS, typename T::Smem* SM, const StepCtx& X) {
  const Go2EnvConfig* C = X.cs;
  const Go2EnvConfig* CT = X.cfg; (void)CT;
  GO2_WIDE_COLD {
    if (lane == 0) {
      int bad = 0;
      if (C->state_guard) {
        for (int k = 0; k < 13; ++k) bad |= !(fabsf(S.root[k]) <= 3.0e38f);
        for (int k = 0; k < GO2_NUM_DOF; ++k) bad |= !(fabsf(S.q[k]) <= 3.0e38f) | !(fabsf(S.qd[k]) <= 3.0e38f);
        if (bad) {
          for (int k = 0; k < 3; ++k) S.root[k] = S.env_origin[k] + C->base_init_state[k];
          for (int k = 3; k < 7; ++k) S.root[k] = C->base_init_state[k];
          for (int k = 7; k < 13; ++k) S.root[k] = 0.0f;
          for (int k = 0; k < GO2_NUM_DOF; ++k) { S.q[k] = C->default_dof_pos[k]; S.qd[k] = 0.0f; S.lqd[k] = 0.0f; S.tq[k] = 0.0f; }
          for (int b = 0; b < GO2_NUM_REPORT; ++b) { S.cf[b][0] = 0.0f; S.cf[b][1] = 0.0f; S.cf[b][2] = 0.0f; }
        }
      }
      S.bad = bad;
    }
  } GO2_SYNC_WARP();
}

// feet position / velocity at the current configuration (rigid_body_states refresh, legged_robot.py:109)
template <class T>
GO2_HD void feet_kinematics(GO2_LANE_ARGS, typename T::Smem* SM, const StepCtx& X) {
  const Go2Model* M = X.mdl;
  GO2_WIDE_COLD {   // lanes 0..3 of the env's own warp: once per step, between WIDE phases (no CTA barrier on either side)
    if (lane < 4) {
      M3 Rw = quat_to_mat(S.root[3], S.root[4], S.root[5], S.root[6]);
      V3 pw = ld3(S.root);
      V6 v;
      v.a = mulT(Rw, ld3(S.root + 10)); v.l = mulT(Rw, ld3(S.root + 7));
      for (int i = 0; i < 3; ++i) {
        const int j = 3 * lane + i, ax = (i == 0) ? 0 : 1;
        float q = S.q[j], sn = sinf(q), cn = cosf(q);
        V3 r = ld3(M->joint_origin[j]);
        pw = pw + mul(Rw, r);
        M3 Rn = Rw;
        const int ci = (ax + 1) % 3, cj = (ax + 2) % 3;
        for (int k = 0; k < 3; ++k) {
          float a = Rw.m[3 * k + ci], b = Rw.m[3 * k + cj];
          Rn.m[3 * k + ci] = cn * a + sn * b; Rn.m[3 * k + cj] = -sn * a + cn * b;
        }
        Rw = Rn;
        V3 wl = v.a, vl = v.l + cross(v.a, r);
        if (ax == 0) { v.a = rot_inv<0>(cn, sn, wl); v.l = rot_inv<0>(cn, sn, vl); }
        else { v.a = rot_inv<1>(cn, sn, wl); v.l = rot_inv<1>(cn, sn, vl); }
        addcomp(v.a, ax, S.qd[j]);
      }
      V3 r = ld3(M->foot_offset[lane]);
      st3(S.feet[lane], pw + mul(Rw, r));
      st3(S.feet[lane] + 3, mul(Rw, v.l + cross(v.a, r)));
    }
  } GO2_SYNC_WARP();
}

// ================================================================================================ commands / reset (lane 0)
// legged_robot.py:423-592 for GO2Cfg (dynamic_resample_commands, no heading command); isaacgym_utils.py:32-55
template <class SMT>
GO2_HD void resample_commands(SMT& S, const StepCtx& X, int e, int stream) {
  const Go2EnvConfig* C = X.cs;
  const Go2EnvConfig* CT = X.cfg; (void)CT;
  const Go2StepParams* sp = X.sp;
  const uint32_t ge = (uint32_t)(C->env_offset + e);
  U4 r0 = philox(ge, sp->common_step_counter, (uint32_t)stream, 0, C->seed_lo, C->seed_hi);
  U4 r1 = philox(ge, sp->common_step_counter, (uint32_t)stream, 1, C->seed_lo, C->seed_hi);
  const float* rng = S.cmd_rng;
  float ep_len = (float)S.ep_len, max_len = (float)C->max_episode_length;
  float accn = sqrtf(S.acc_xy[0] * S.acc_xy[0] + S.acc_xy[1] * S.acc_xy[1]);
  float remaining = fmaxf(GO2_FADD(GO2_FMUL(0.625f, C->terrain_length), -GO2_FMUL(accn, C->resampling_time)), 0.0f);
  const float full = C->resampling_time / C->dt;
  S.resamp_step = full;
  const bool heading = C->heading_command != 0;
  S.stop_heading = 0;                                                   // legged_robot.py:431
  if (C->dynamic_resample_commands) {
    float vlow = fmaxf(remaining / GO2_FMUL(GO2_FADD(max_len - ep_len, 1e-9f), C->dt), 0.0f);
    for (int a = 0; a < 2; ++a) {
      float lo = rng[2 * a], hi = rng[2 * a + 1];
      float wneg = fmaxf(-vlow - lo, 0.0f), wpos = fmaxf(hi - vlow, 0.0f);
      float total = GO2_FADD(GO2_FADD(wneg, wpos), 1e-6f);
      float u = GO2_FMUL(u01(a == 0 ? r0.x : r0.y), total);
      S.cmd[a] = (u < wneg) ? GO2_FADD(lo, u) : GO2_FADD(hi - wpos, u - wneg);
    }
    if (heading) S.cmd[3] = affine(S.hrng[1] - S.hrng[0], u01(r0.z), S.hrng[0]);       // the same draw feeds the heading target (:468-472)
    else
    S.cmd[2] = affine(rng[5] - rng[4], u01(r0.z), rng[4]);
  } else {
    S.cmd[0] = GO2_FADD(rng[0], GO2_FMUL(u01(r0.x), rng[1] - rng[0]));
    S.cmd[1] = GO2_FADD(rng[2], GO2_FMUL(u01(r0.y), rng[3] - rng[2]));
    if (heading) S.cmd[3] = GO2_FADD(S.hrng[0], GO2_FMUL(u01(r0.z), S.hrng[1] - S.hrng[0]));
    else
    S.cmd[2] = GO2_FADD(rng[4], GO2_FMUL(u01(r0.z), rng[5] - rng[4]));
    float nrm = sqrtf(S.cmd[0] * S.cmd[0] + S.cmd[1] * S.cmd[1]);
    if (!(nrm > 0.2f)) { S.cmd[0] = 0; S.cmd[1] = 0; }
  }
  float prob = u01(r0.w), min_p = 0, max_p = 0;
  if (C->limit_vel_prob > 0) {
    max_p += C->limit_vel_prob;
    bool lim = prob >= min_p && prob < max_p;
    if (lim) {
      bool change = true;
      if (C->limit_vel_invert_when_continuous && S.last_lim) {
        S.cmd[0] *= -1.0f; S.cmd[1] *= -1.0f; S.cmd[2] *= -1.0f;
        change = false;
      }
      if (change) {  // limit_vel_comb = product([-1,1],[-1,1],[-1,0,1]), legged_robot.py:827-831
        int idx = (int)(r1.x % 12u), cx = idx / 6, cy = (idx / 3) % 2, cz = idx % 3;
        S.cmd[0] = cx == 0 ? rng[0] : rng[1];
        S.cmd[1] = cy == 0 ? rng[2] : rng[3];
        S.cmd[2] = cz == 0 ? rng[4] : (cz == 1 ? 0.0f : rng[5]);
      }
      if (heading && C->stop_heading_at_limit) S.stop_heading = 1;      // :547-548
    }
    S.last_lim = lim ? 1 : 0;
    min_p += C->limit_vel_prob;
  }
  if (sp->zero_command_proba > 0) {
    max_p += sp->zero_command_proba;
    float next = (max_len - ep_len) - remaining / GO2_FADD(GO2_FMUL(GO2_FMUL(0.8f, sp->max_lin_vel), C->dt), 1e-9f);
    next = fminf(fmaxf(next, 0.0f), full);
    if (prob >= min_p && prob < max_p && next > 0) {
      S.cmd[0] = 0; S.cmd[1] = 0;
      S.resamp_step = next;
      if (C->limit_ang_vel_at_zero_command_prob > 0 && u01(r1.y) < C->limit_ang_vel_at_zero_command_prob) {
        S.cmd[2] = (u01(r1.z) < 0.5f) ? rng[4] : rng[5];
        if (heading) S.stop_heading = 1;                                // :581-582
      }
    }
  }
  if (C->turn_over && GO2_EXT_PTR(const float*, C, ext_turn_over_timer)[e] > 0.0f) {      // turn-over zero-command time (legged_robot.py:585-590)
    S.cmd[0] = 0; S.cmd[1] = 0; S.cmd[2] = 0;
    S.stop_heading = 1;
  }
  S.acc_xy[0] += S.cmd[0]; S.acc_xy[1] += S.cmd[1];
}

// yaw-rate command from the heading target (legged_robot.py:411-419; quat_apply and wrap_to_pi in torch's operation order)
template <class SMT>
GO2_HD void heading_to_yaw(SMT& S) {
  const float qx = S.root[3], qy = S.root[4], qz = S.root[5], qw = S.root[6];
  // t = 2 (q_xyz x [1,0,0]) = 2 (0, qz, -qy);  forward = [1,0,0] + qw t + q_xyz x t
  const float ty = GO2_FMUL(qz, 2.0f), tz = GO2_FMUL(-qy, 2.0f);
  const float fx = GO2_FADD(1.0f, GO2_FADD(GO2_FMUL(qy, tz), -GO2_FMUL(qz, ty)));
  const float fy = GO2_FADD(GO2_FMUL(qw, ty), GO2_FADD(GO2_FMUL(qz, 0.0f), -GO2_FMUL(qx, tz)));
  const float hd = atan2f(fy, fx);
  float a = fmodf(S.cmd[3] - hd, 6.2831855f);                           // torch: angles %= 2 pi (result takes the divisor's sign)
  if (a != 0.0f && a < 0.0f) a = GO2_FADD(a, 6.2831855f);
  if (a > 3.1415927f) a = GO2_FADD(a, -6.2831855f);
  S.cmd[2] = fminf(fmaxf(GO2_FMUL(0.5f, a), S.cmd_rng[4]), S.cmd_rng[5]);
}

// The reward functions that are inactive in every registered go2 task (legged_robot.py:1236-1441, go2_env.py:62-68; enum Go2XReward), evaluated on
// lane 0 when the config gives one of them a non-zero scale (Go2EnvConfig.num_xrew > 0).  A function the reference does not call (zero scale) does
// not advance its state (feet_air_time / last_contacts, last_contacts2).  rew receives the scaled terms; the termination term is returned
// separately: the reference adds it after the only_positive_rewards clip (legged_robot.py:268-272).
template <class SMT>
GO2_HD void extra_rewards(SMT& S, const StepCtx& X, int e, bool need_to, float& rew, float& term_rew) {
  const Go2EnvConfig* C = X.cs;
  const Go2EnvConfig* CT = X.cfg;
  const Go2Model* M = X.mdl;
  float* sums = GO2_EXT_PTR(float*, C, ext_xrew_sums) + (size_t)e * GO2_NUM_XREW;
  float* st = GO2_EXT_PTR(float*, C, ext_xrew_state) + (size_t)e * 12;
  const float* sc = C->xrew_scales;
  const float* tsc = C->to_xscales;      // turn_over scales: a term is evaluated when either scale is set (legged_robot.py:925-930)
  float tv[GO2_NUM_XREW];
  for (int k = 0; k < GO2_NUM_XREW; ++k) tv[k] = 0.0f;
  const float cmd_xy = sqrtf(S.cmd[0] * S.cmd[0] + S.cmd[1] * S.cmd[1]);
  bool contact[4];
  for (int l = 0; l < 4; ++l) contact[l] = S.cf[6 + 4 * l][2] > 1.0f;                       // feet_indices = reported bodies 6 + 4 l
  tv[GO2_XREW_ORIENTATION] = S.pg[0] * S.pg[0] + S.pg[1] * S.pg[1];                          // :1236-1238
  if (sc[GO2_XREW_BASE_HEIGHT] != 0.0f || tsc[GO2_XREW_BASE_HEIGHT] != 0.0f) {                                                    // :1245-1259
    float nfc = 0.0f, fcp[3] = {0.0f, 0.0f, 0.0f};
    for (int l = 0; l < 4; ++l) {
      const bool filt = contact[l] || st[8 + l] != 0.0f;
      st[8 + l] = contact[l] ? 1.0f : 0.0f;
      if (filt) { nfc += 1.0f; for (int k = 0; k < 3; ++k) fcp[k] += S.feet[l][k]; }
    }
    const float den = fmaxf(nfc, 1.0f);
    float bh = 0.0f;
    for (int k = 0; k < 3; ++k) bh += (fcp[k] / den - S.root[k]) * S.pg[k];
    tv[GO2_XREW_BASE_HEIGHT] = (bh - C->base_height_target) * (bh - C->base_height_target) * (nfc > 0.0f ? 1.0f : 0.0f);
  }
  float s_qd = 0.0f, s_vl = 0.0f, s_tl = 0.0f, s_def = 0.0f;
  for (int j = 0; j < GO2_NUM_DOF; ++j) {
    s_qd += S.qd[j] * S.qd[j];                                                               // dof_vel :1265-1267
    s_vl += fminf(fmaxf(fabsf(S.qd[j]) - M->vel_limit[j] * C->soft_dof_vel_limit, 0.0f), 1.0f);   // dof_vel_limits :1291-1294
    s_tl += fmaxf(fabsf(S.tq[j]) - M->effort[j] * C->soft_torque_limit, 0.0f);              // torque_limits :1296-1298
    s_def += fabsf(S.q[j] - CT->default_dof_pos[j]);                                         // similar_to_default :1416-1418
  }
  tv[GO2_XREW_DOF_VEL] = s_qd; tv[GO2_XREW_DOF_VEL_LIMITS] = s_vl; tv[GO2_XREW_TORQUE_LIMITS] = s_tl; tv[GO2_XREW_SIMILAR_TO_DEFAULT] = s_def;
  tv[GO2_XREW_TERMINATION] = (S.reset && !S.tout) ? 1.0f : 0.0f;                             // :1281-1283
  if (sc[GO2_XREW_FEET_AIR_TIME] != 0.0f || tsc[GO2_XREW_FEET_AIR_TIME] != 0.0f) {                                                  // :1347-1358
    float r = 0.0f;
    for (int l = 0; l < 4; ++l) {
      const bool filt = contact[l] || st[4 + l] != 0.0f;
      st[4 + l] = contact[l] ? 1.0f : 0.0f;
      const bool first = st[l] > 0.0f && filt;
      st[l] += C->dt;
      r += (st[l] - 0.5f) * (first ? 1.0f : 0.0f);
      if (filt) st[l] = 0.0f;
    }
    tv[GO2_XREW_FEET_AIR_TIME] = r * (cmd_xy > 0.1f ? 1.0f : 0.0f);
  }
  {
    bool stumble = false;                                                                    // :1360-1363
    float s_fc = 0.0f;                                                                       // feet_contact_forces :1369-1371
    for (int l = 0; l < 4; ++l) {
      const float* f = S.cf[6 + 4 * l];
      stumble = stumble || (sqrtf(f[0] * f[0] + f[1] * f[1]) > 5.0f * fabsf(f[2]));
      s_fc += fmaxf(sqrtf(f[0] * f[0] + f[1] * f[1] + f[2] * f[2]) - C->max_contact_force, 0.0f);
    }
    tv[GO2_XREW_STUMBLE] = stumble ? 1.0f : 0.0f;
    tv[GO2_XREW_FEET_CONTACT_FORCES] = s_fc;
  }
  tv[GO2_XREW_STAND_STILL] = s_def * (cmd_xy < 0.1f ? 1.0f : 0.0f);                          // :1365-1367
  tv[GO2_XREW_UPRIGHT] = (-1.0f - S.pg[2]) / 2.0f;                                           // :1420-1421
  {                                                                                          // legs_distance :1423-1441
    float ly[4];
    for (int l = 0; l < 4; ++l) ly[l] = quat_rotate_inverse(S.root + 3, mk(S.feet[l][0] - S.root[0], S.feet[l][1] - S.root[1], S.feet[l][2] - S.root[2])).y;
    const float df = fmaxf(C->min_legs_distance - (ly[0] - ly[1]), 0.0f), dr = fmaxf(C->min_legs_distance - (ly[2] - ly[3]), 0.0f);
    tv[GO2_XREW_LEGS_DISTANCE] = df * df + dr * dr;
  }
  if (sc[GO2_XREW_X_COMMAND_HIP_REGULAR] != 0.0f || tsc[GO2_XREW_X_COMMAND_HIP_REGULAR] != 0.0f) {                                          // go2_env.py:62-68 (0 / 0 at an all-zero command, like the reference)
    const float ratio = fabsf(S.cmd[0]) / sqrtf(S.cmd[0] * S.cmd[0] + S.cmd[1] * S.cmd[1] + S.cmd[2] * S.cmd[2]);
    tv[GO2_XREW_X_COMMAND_HIP_REGULAR] = (fabsf(S.q[0] + S.q[3]) + fabsf(S.q[6] + S.q[9])) * ratio;
  }
  term_rew = 0.0f;
  for (int k = 0; k < GO2_NUM_XREW; ++k) {
    if (sc[k] == 0.0f && tsc[k] == 0.0f) continue;
    const float rk = tv[k] * ((need_to && k != GO2_XREW_TERMINATION) ? tsc[k] : sc[k]) * X.sp->xrew_curriculum[k];
    if (k == GO2_XREW_TERMINATION) term_rew = rk; else rew += rk;
    sums[k] += rk;
  }
}

#if defined(__CUDACC__)
#define GO2_ATOMIC_ADD(p, v) atomicAdd((p), (v))
#define GO2_ATOMIC_ADD_FIXED(acc, k, v) atomicAdd(reinterpret_cast<unsigned long long*>((acc) + GO2_EP_ACC_FIXED_OFF) + (k), (unsigned long long)__float2ll_rn((v) * GO2_EP_FIXED_ONE))
#define GO2_ATOMIC_ADD_FIXED64(acc64, k, v) atomicAdd(reinterpret_cast<unsigned long long*>(acc64) + (k), (unsigned long long)__float2ll_rn((v) * GO2_EP_FIXED_ONE))
#else
#define GO2_ATOMIC_ADD(p, v) (*(p) += (v))
#define GO2_ATOMIC_ADD_FIXED(acc, k, v) (reinterpret_cast<long long*>((acc) + GO2_EP_ACC_FIXED_OFF)[k] += llrintf((v) * GO2_EP_FIXED_ONE))
#define GO2_ATOMIC_ADD_FIXED64(acc64, k, v) ((acc64)[k] += llrintf((v) * GO2_EP_FIXED_ONE))
#endif

// reset_idx for this env (legged_robot.py:180-245, :620-707, :1143-1169); `initial` = the reset at construction
template <class T>
GO2_HD void reset_phases(GO2_LANE_ARGS, typename T::Smem* SM, const StepCtx& X, bool initial) {
  const Go2EnvConfig* C = X.cs;
  const Go2EnvConfig* CT = X.cfg; (void)CT;
  const Go2EnvBuffers* B = X.buf;
  const Go2StepParams* sp = X.sp;
  GO2_WIDE_COLD if (initial || S.reset) {   // only the envs of the group that reset
    const uint32_t ge = (uint32_t)(C->env_offset + e);
    if (lane < GO2_NUM_DOF) {
      const size_t o = (size_t)e * GO2_NUM_DOF + lane;
      U4 r = philox(ge, sp->common_step_counter, ST_RESET_DR, (uint32_t)lane, C->seed_lo, C->seed_hi);
      if (C->randomize_motor_strength) B->motor_strengths[o] = affine(C->motor_strength_range[1], u01(r.x), C->motor_strength_range[0]);
      if (C->randomize_motor_zero_offset) B->motor_zero_offsets[o] = affine(C->motor_zero_offset_range[1], u01(r.y), C->motor_zero_offset_range[0]);
      if (C->randomize_pd_gains) {
        B->p_gains_multiplier[o] = affine(C->kp_mult_range[1], u01(r.z), C->kp_mult_range[0]);
        B->d_gains_multiplier[o] = affine(C->kd_mult_range[1], u01(r.w), C->kd_mult_range[0]);
      }
      U4 rq = philox(ge, sp->common_step_counter, ST_RESET_STATE, (uint32_t)(lane / 4), C->seed_lo, C->seed_hi);
      S.q[lane] = GO2_FMUL(CT->default_dof_pos[lane], GO2_FADD(u01(pick(rq, lane % 4)), 0.5f));
      S.qd[lane] = 0; S.act[lane] = 0; S.lact[lane] = 0; S.lqd[lane] = 0;
    }
    if (lane >= 16 && lane < 16 + GO2_NUM_REW) {
      const int k = lane - 16;
      float* es = B->episode_sums + (size_t)e * GO2_NUM_REW + k;
      if (!initial) GO2_ATOMIC_ADD_FIXED(B->ep_accum, k, *es);      // 2^-20 fixed point, integer atomics: the sum does not depend on the order
      *es = 0;
    }
    if (lane == 30 && !initial) GO2_ATOMIC_ADD(B->ep_accum + GO2_NUM_REW + 10, 1.0f);
    if (C->num_xrew > 0 && lane >= 16 && lane < 16 + GO2_NUM_XREW) {     // the extra terms' episode sums (legged_robot.py:229-233), feet_air_time (:220)
      const int k = lane - 16;
      float* xs = GO2_EXT_PTR(float*, C, ext_xrew_sums) + (size_t)e * GO2_NUM_XREW + k;
      if (!initial) GO2_ATOMIC_ADD_FIXED64(GO2_EXT_PTR(long long*, C, ext_xrew_log), k, *xs);
      *xs = 0;
      if (k < 4) GO2_EXT_PTR(float*, C, ext_xrew_state)[(size_t)e * 12 + k] = 0;
    }
    if (lane == 0) {
      U4 s3 = philox(ge, sp->common_step_counter, ST_RESET_STATE, 3, C->seed_lo, C->seed_hi);
      U4 s4 = philox(ge, sp->common_step_counter, ST_RESET_STATE, 4, C->seed_lo, C->seed_hi);
      U4 s5 = philox(ge, sp->common_step_counter, ST_RESET_STATE, 5, C->seed_lo, C->seed_hi);
      if (C->terrain_curriculum && !initial && C->mesh_type != 0) {  // _update_terrain_curriculum
        float dist = S.max_move;
        float accn = sqrtf(S.acc_xy[0] * S.acc_xy[0] + S.acc_xy[1] * S.acc_xy[1]);
        bool up = dist > C->terrain_length / 2, down;
        if (C->move_down_by_accumulated_xy_command)
          down = (dist < GO2_FMUL(GO2_FMUL(accn, GO2_FMUL(C->resampling_time, 1 - sp->zero_command_proba)), 0.5f)) && !up;
        else
          down = (dist < GO2_FMUL(GO2_FMUL(sqrtf(S.cmd[0] * S.cmd[0] + S.cmd[1] * S.cmd[1]), C->max_episode_length_s), 0.5f)) && !up;
        int lvl = S.level + (up ? 1 : 0) - (down ? 1 : 0);
        if (lvl >= C->num_levels) lvl = (int)(s3.w % (uint32_t)C->num_levels);
        else lvl = max(lvl, 0);
        S.level = lvl;
        const float* org = B->terrain_origins + ((size_t)lvl * C->num_types + S.ttype) * 3;
        S.env_origin[0] = GO2_LDG(org); S.env_origin[1] = GO2_LDG(org + 1); S.env_origin[2] = GO2_LDG(org + 2);
        S.max_move = 0;
      }
      const float PI_F = 3.14159265358979323846f;
      float yaw = affine(2 * PI_F, u01(s3.x), -PI_F);
      for (int i = 0; i < 13; ++i) S.root[i] = C->base_init_state[i];
      S.root[3] = 0; S.root[4] = 0; S.root[5] = sinf(GO2_FMUL(yaw, 0.5f)); S.root[6] = cosf(GO2_FMUL(yaw, 0.5f));
      if (C->turn_over) {       // flipped initial poses (legged_robot.py:642-691): on the back (roll pi), on a side (roll +- pi / 2) or upright
        float* tt = GO2_EXT_PTR(float*, C, ext_turn_over_timer) + e;
        *tt = 0.0f;
        const float rp = u01(s5.z);
        const float* pr = C->turn_over_proportions;       // cumulative: [back, back + side, back + side + none]
        const bool back = rp >= 0.0f && rp < pr[0], side = rp >= pr[0] && rp < pr[1], none = rp >= pr[1] && rp < pr[2];
        if (back || side) {
          U4 s6 = philox(ge, sp->common_step_counter, ST_RESET_STATE, 6, C->seed_lo, C->seed_hi);
          float roll;
          if (back) { S.root[2] = affine(C->turn_over_back_height[1], u01(s6.x), C->turn_over_back_height[0]); roll = PI_F; *tt = C->turn_over_zero_time_back; }
          else { S.root[2] = affine(C->turn_over_side_height[1], u01(s6.y), C->turn_over_side_height[0]); roll = (u01(s5.w) < 0.5f) ? 0.5f * PI_F : -0.5f * PI_F; *tt = C->turn_over_zero_time_side; }
          // quat_from_euler_xyz(roll, 0, yaw) in torch's operation order (cp = 1, sp = 0)
          const float cy = cosf(GO2_FMUL(yaw, 0.5f)), sy = sinf(GO2_FMUL(yaw, 0.5f)), cr = cosf(GO2_FMUL(roll, 0.5f)), sr = sinf(GO2_FMUL(roll, 0.5f));
          S.root[6] = GO2_FADD(GO2_FMUL(GO2_FMUL(cy, cr), 1.0f), GO2_FMUL(GO2_FMUL(sy, sr), 0.0f));
          S.root[3] = GO2_FADD(GO2_FMUL(GO2_FMUL(cy, sr), 1.0f), -GO2_FMUL(GO2_FMUL(sy, cr), 0.0f));
          S.root[4] = GO2_FADD(GO2_FMUL(GO2_FMUL(cy, cr), 0.0f), GO2_FMUL(GO2_FMUL(sy, sr), 1.0f));
          S.root[5] = GO2_FADD(GO2_FMUL(GO2_FMUL(sy, cr), 1.0f), -GO2_FMUL(GO2_FMUL(cy, sr), 0.0f));
        } else if (!none) {     // proportions that do not add up to one leave the configured orientation (base_init_state) for the remainder
          for (int i = 3; i < 7; ++i) S.root[i] = C->base_init_state[i];
        }
      }
      for (int i = 0; i < 3; ++i) S.root[i] = GO2_FADD(S.root[i], S.env_origin[i]);
      if (C->custom_origins) { S.root[0] = GO2_FADD(S.root[0], affine(2.0f, u01(s3.y), -1.0f)); S.root[1] = GO2_FADD(S.root[1], affine(2.0f, u01(s3.z), -1.0f)); }
      S.root[7] = u01(s4.x) - 0.5f; S.root[8] = u01(s4.y) - 0.5f; S.root[9] = u01(s4.z) - 0.5f;
      S.root[10] = u01(s4.w) - 0.5f; S.root[11] = u01(s5.x) - 0.5f; S.root[12] = u01(s5.y) - 0.5f;
      S.ep_len = 0;
      S.reset = 1;
      S.acc_xy[0] = 0; S.acc_xy[1] = 0;
      resample_commands(S, X, e, ST_CMD_RESET);
    }
  } GO2_SYNC_WARP();
}

// ================================================================================================ load / store
// predicated global loads: every lane issues the same load instructions (no branch between them), so all of a thread's loads are in flight
// together and load_env costs ONE memory round trip instead of one per `if (lane == ...)` block
#if defined(__CUDACC__)
__device__ __forceinline__ float ldg_if(bool p, const float* a) {
  float v = 0.0f;
  asm volatile("{\n .reg .pred q;\n setp.ne.b32 q, %2, 0;\n @q ld.global.f32 %0, [%1];\n}" : "+f"(v) : "l"(a), "r"((int)p));
  return v;
}
__device__ __forceinline__ int ldg_if(bool p, const int32_t* a) {
  int v = 0;
  asm volatile("{\n .reg .pred q;\n setp.ne.b32 q, %2, 0;\n @q ld.global.b32 %0, [%1];\n}" : "+r"(v) : "l"(a), "r"((int)p));
  return v;
}
__device__ __forceinline__ int ldg_if(bool p, const uint8_t* a) {
  int v = 0;
  asm volatile("{\n .reg .pred q;\n setp.ne.b32 q, %2, 0;\n @q ld.global.u8 %0, [%1];\n}" : "+r"(v) : "l"(a), "r"((int)p));
  return v;
}
#else
static inline float ldg_if(bool p, const float* a) { return p ? *a : 0.0f; }
static inline int ldg_if(bool p, const int32_t* a) { return p ? *a : 0; }
static inline int ldg_if(bool p, const uint8_t* a) { return p ? (int)*a : 0; }
#endif

template <class T>
GO2_HD void load_env(GO2_LANE_ARGS, typename T::Smem* SM, const StepCtx& X) {
  const Go2EnvConfig* C = X.cs;
  const Go2EnvConfig* CT = X.cfg; (void)CT;
  const Go2EnvBuffers* B = X.buf;
  GO2_WIDE {
    // ---- every global load of the lane, issued back to back (the buffers are read-only for the duration of the loads: the kernel writes
    // them in store_state / the post-physics phases only)
    constexpr int NI = GO2_NUM_DYN * GO2_INERTIA_STRIDE, NCF = GO2_NUM_REPORT * 3;
    const bool pj = lane < GO2_NUM_DOF;
    const size_t o = (size_t)e * GO2_NUM_DOF + (pj ? lane : 0);
    float vin[(NI + 31) / 32], vcf[(NCF + 31) / 32];
    GO2_UNROLL for (int k = 0; k < (NI + 31) / 32; ++k) vin[k] = ldg_if(lane + 32 * k < NI, B->body_inertia + (size_t)e * NI + min(lane + 32 * k, NI - 1));
    GO2_UNROLL for (int k = 0; k < (NCF + 31) / 32; ++k) vcf[k] = ldg_if(lane + 32 * k < NCF, B->contact_forces + (size_t)e * NCF + min(lane + 32 * k, NCF - 1));
    const float v_root = ldg_if(lane < 13, B->root_states + (size_t)e * 13 + min(lane, 12));
    const float v_q = ldg_if(pj, B->dof_pos + o), v_qd = ldg_if(pj, B->dof_vel + o), v_la = ldg_if(pj, B->last_actions + o);
    const float v_lla = ldg_if(pj, B->last_last_actions + o), v_lqd = ldg_if(pj, B->last_dof_vel + o), v_tq = ldg_if(pj, B->torques + o);
    const float v_a = ldg_if(pj, (X.actions_in ? X.actions_in : B->actions) + o);
    const float v_pg = ldg_if(pj, B->p_gains_multiplier + o), v_dg = ldg_if(pj, B->d_gains_multiplier + o);
    const float v_mzo = ldg_if(pj, B->motor_zero_offsets + o), v_ms = ldg_if(pj, B->motor_strengths + o);
    const float v_kp = ldg_if(pj, CT->kp + (pj ? lane : 0)), v_kd = ldg_if(pj, CT->kd + (pj ? lane : 0));
    const float v_ddp = ldg_if(pj, CT->default_dof_pos + (pj ? lane : 0)), v_eff = ldg_if(pj, X.mdl->effort + (pj ? lane : 0));
    const float v_qlo = ldg_if(pj, X.mdl->q_lower + (pj ? lane : 0)), v_qhi = ldg_if(pj, X.mdl->q_upper + (pj ? lane : 0));
    const float v_vl = ldg_if(pj, X.mdl->vel_limit + (pj ? lane : 0));
    const bool p_cmd = lane >= 12 && lane < 16, p_rng = lane >= 16 && lane < 22, p_org = lane >= 22 && lane < 25;
    const float v_cmd = ldg_if(p_cmd, B->commands + (size_t)e * GO2_NUM_CMD + (p_cmd ? lane - 12 : 0));
    const float v_rng = ldg_if(p_rng, B->env_command_ranges + (size_t)e * 6 + (p_rng ? lane - 16 : 0));
    const float v_org = ldg_if(p_org, B->env_origins + (size_t)e * 3 + (p_org ? lane - 22 : 0));
    const int v_eplen = ldg_if(lane == 25, B->episode_length_buf + e);
    const float v_rs = ldg_if(lane == 25, B->commands_resampling_step + e);
    const float v_ax = ldg_if(lane == 26, B->commands_xy_accumulation + (size_t)e * 2), v_ay = ldg_if(lane == 26, B->commands_xy_accumulation + (size_t)e * 2 + 1);
    const float v_mm = ldg_if(lane == 27, B->max_move_distance + e);
    const int v_ll = ldg_if(lane == 27, B->last_is_limit_vel + e);
    const int v_lvl = ldg_if(lane == 28, B->terrain_levels + e), v_tt = ldg_if(lane == 28, B->terrain_types + e), v_tid = ldg_if(lane == 28, B->terrain_ids + e);
    const float v_fr = ldg_if(lane == 30, B->friction_coeffs + e), v_re = ldg_if(lane == 30, B->restitutions + e);
    // ---- shared-memory writes
    GO2_UNROLL for (int k = 0; k < (NI + 31) / 32; ++k) if (lane + 32 * k < NI) S.inertia[lane + 32 * k] = vin[k];
    GO2_UNROLL for (int k = 0; k < (NCF + 31) / 32; ++k) if (lane + 32 * k < NCF) S.cf[(lane + 32 * k) / 3][(lane + 32 * k) % 3] = vcf[k];
    if (lane < 13) S.root[lane] = v_root;
    if (pj) {
      S.q[lane] = v_q; S.qd[lane] = v_qd;
      S.lact[lane] = v_la; S.llact[lane] = v_lla; S.lqd[lane] = v_lqd;
      S.tq[lane] = v_tq;
      S.act[lane] = fminf(fmaxf(v_a, -C->clip_actions), C->clip_actions);
      L.kp = v_kp * v_pg; L.kd = v_kd * v_dg;
      L.mzo = v_mzo; L.mstr = v_ms;
      L.ddp = v_ddp; L.eff = v_eff; L.qlo = v_qlo; L.qhi = v_qhi; L.vlim = v_vl;
    }
    if (p_cmd) S.cmd[lane - 12] = v_cmd;
    if (p_rng) S.cmd_rng[lane - 16] = v_rng;
    if (p_org) S.env_origin[lane - 22] = v_org;
    if (lane == 25) { S.ep_len = v_eplen; S.resamp_step = v_rs; }
    if (lane == 26) { S.acc_xy[0] = v_ax; S.acc_xy[1] = v_ay; }
    if (lane == 27) { S.max_move = v_mm; S.last_lim = v_ll; }
    if (lane == 28) { S.level = v_lvl; S.ttype = v_tt; S.tid = v_tid; }
    if (lane == 30) {
      S.stop_heading = 0; S.hrng[0] = 0.0f; S.hrng[1] = 0.0f;
      if (C->heading_command) {
        const float* hr = GO2_EXT_PTR(const float*, C, ext_heading_ranges);
        S.stop_heading = GO2_EXT_PTR(const uint8_t*, C, ext_stop_heading)[e]; S.hrng[0] = hr[(size_t)e * 2]; S.hrng[1] = hr[(size_t)e * 2 + 1];
      } else if (C->turn_over) S.stop_heading = GO2_EXT_PTR(const uint8_t*, C, ext_stop_heading)[e];    // the flag is also set by the turn-over zero-command time
      S.mu_env = 0.5f * (C->terrain_friction + v_fr);
      S.rest_env = 0.5f * (C->terrain_restitution + v_re);
    }
    if (lane == 29) {
      S.delay_start = 0;
      if (C->randomize_action_delay && X.sp)
        S.delay_start = (int)(philox((uint32_t)(C->env_offset + e), X.sp->common_step_counter, ST_DELAY, 0, C->seed_lo, C->seed_hi).x % (uint32_t)(C->decimation + 1));
    }
  } GO2_SYNC_WARP();
}

template <class T>
GO2_HD void store_state(GO2_LANE_ARGS, typename T::Smem* SM, const StepCtx& X) {
  const Go2EnvBuffers* B = X.buf;
  GO2_WIDE_COLD {
    if (lane < 13) B->root_states[(size_t)e * 13 + lane] = S.root[lane];
    if (lane < GO2_NUM_DOF) {
      const size_t o = (size_t)e * GO2_NUM_DOF + lane;
      B->dof_pos[o] = S.q[lane]; B->dof_vel[o] = S.qd[lane]; B->torques[o] = S.tq[lane];
      B->actions[o] = S.act[lane]; B->last_actions[o] = S.lact[lane]; B->last_last_actions[o] = S.llact[lane];
      B->last_dof_vel[o] = S.lqd[lane];
    }
    if (lane >= 12 && lane < 16) B->commands[(size_t)e * GO2_NUM_CMD + lane - 12] = S.cmd[lane - 12];
    if (lane >= 22 && lane < 25) B->env_origins[(size_t)e * 3 + lane - 22] = S.env_origin[lane - 22];
    if (lane == 25) { B->episode_length_buf[e] = S.ep_len; B->commands_resampling_step[e] = S.resamp_step; }
    if (lane == 26) { B->commands_xy_accumulation[(size_t)e * 2] = S.acc_xy[0]; B->commands_xy_accumulation[(size_t)e * 2 + 1] = S.acc_xy[1]; }
    if (lane == 27) { B->max_move_distance[e] = S.max_move; B->last_is_limit_vel[e] = (uint8_t)S.last_lim; }
    if (lane == 30 && (X.cs->heading_command || X.cs->turn_over)) GO2_EXT_PTR(uint8_t*, X.cs, ext_stop_heading)[e] = (uint8_t)S.stop_heading;
    if (lane == 28) B->terrain_levels[e] = S.level;
    if (lane == 29) { B->reset_buf[e] = (uint8_t)S.reset; B->time_out_buf[e] = (uint8_t)S.tout; }
    GO2_STRIDED(i, GO2_NUM_REPORT * 3) B->contact_forces[(size_t)e * GO2_NUM_REPORT * 3 + i] = S.cf[i / 3][i % 3];
    if (lane < 24) {
      const int l = lane / 6, k = lane % 6;
      if (k < 3) B->feet_pos[(size_t)e * 12 + l * 3 + k] = S.feet[l][k];
      else B->feet_vel[(size_t)e * 12 + l * 3 + (k - 3)] = S.feet[l][k];
    }
  } GO2_SYNC_WARP();
}

// torques for the current substep (legged_robot.py:74-81, :594-618, control_type 'P')
template <class T>
GO2_HD void compute_torques(GO2_LANE_ARGS, typename T::Smem* SM, const StepCtx& X, int sub) {
  const Go2EnvConfig* C = X.cs;
  const Go2EnvConfig* CT = X.cfg; (void)CT;
  const Go2Model* M = X.mdl; (void)M;
  GO2_WIDE {
    if (lane < GO2_NUM_DOF) {
      float a_in = (C->randomize_action_delay && sub < S.delay_start) ? S.lact[lane] : S.act[lane];
      float t = L.kp * (a_in * C->action_scale + L.ddp - S.q[lane] + L.mzo) - L.kd * S.qd[lane];
      if (C->control_type == 1) t = L.kp * (a_in * C->action_scale - S.qd[lane]) - L.kd * (S.qd[lane] - S.lqd[lane]) / C->sim_dt;   // 'V', legged_robot.py:612-613
      else if (C->control_type == 2) t = a_in * C->action_scale;                                                                  // 'T', :614-615
      float lim = L.eff;
      t = fminf(fmaxf(t, -lim), lim);
      if (C->randomize_motor_strength) t *= L.mstr;
      S.tq[lane] = t;                               // what the reference reports (legged_robot.py:79-81)
      S.tau[lane] = fminf(fmaxf(t, -lim), lim);     // effort clamp of the actuator (physics spec)
    }
  } GO2_SYNC_WARP();
}

// ================================================================================================ the full step
template <class T>
GO2_HD void step_env(GO2_LANE_ARGS, typename T::Smem* SM, const StepCtx& X) {
  const Go2EnvConfig* C = X.cs;
  const Go2EnvConfig* CT = X.cfg; (void)CT;
  const Go2Model* M = X.mdl; (void)M;
  const Go2EnvBuffers* B = X.buf;
  const Go2StepParams* sp = X.sp;
  load_env<T>(GO2_LANE_PASS, SM, X);
  for (int sub = 0; sub < C->decimation; ++sub) {
    GO2_COARSE_SYNC();
    compute_torques<T>(GO2_LANE_PASS, SM, X, sub);
    physics_substep<T>(GO2_LANE_PASS, SM, X, sub == C->decimation - 1);
  }
  GO2_COARSE_SYNC();
  GO2_TICK();
  state_guard<T>(GO2_LANE_PASS, SM, X); GO2_TICK();
  feet_kinematics<T>(GO2_LANE_PASS, SM, X); GO2_TICK();
  // ---- post_physics_step (legged_robot.py:102-142)
  GO2_WIDE_COLD {
    // height scan (legged_robot.py:1188-1224, math.py:8-12): yaw-only rotation of the body-frame grid
    if (C->mesh_type == 0) {
      GO2_STRIDED(i, GO2_NUM_HEIGHT) S.heights[i] = 0;
    } else {
      float qz = S.root[5], qw = S.root[6];
      float nrm = fmaxf(sqrtf(qz * qz + qw * qw), 1e-9f);
      qz /= nrm; qw /= nrm;
#pragma unroll
      for (int k6 = 0; k6 < (GO2_NUM_HEIGHT + 31) / 32; ++k6) {      // unrolled and branch-free: the 3 x 6 heightfield loads of a lane are all in flight together
        const int i0 = lane + 32 * k6, i = min(i0, GO2_NUM_HEIGHT - 1);  // lanes past the end recompute the last sample and do not store
        float bx = CT->height_points[i][0], by = CT->height_points[i][1];
        float tx = GO2_FMUL(2.0f, -GO2_FMUL(qz, by)), ty = GO2_FMUL(2.0f, GO2_FMUL(qz, bx));
        float px = GO2_FADD(GO2_FADD(bx, GO2_FMUL(qw, tx)), -GO2_FMUL(qz, ty)), py = GO2_FADD(GO2_FADD(by, GO2_FMUL(qw, ty)), GO2_FMUL(qz, tx));
        px = GO2_FADD(GO2_FADD(px, S.root[0]), C->border); py = GO2_FADD(GO2_FADD(py, S.root[1]), C->border);
        int ix = (int)(px / C->hscale), iy = (int)(py / C->hscale);  // .long(): truncation toward zero
        ix = min(max(ix, 0), C->hf_rows - 2); iy = min(max(iy, 0), C->hf_cols - 2);
        const int16_t* p = B->height_samples + (size_t)ix * C->hf_cols + iy;
        int16_t h1 = GO2_LDG(p), h2 = GO2_LDG(p + C->hf_cols), h3 = GO2_LDG(p + 1);
        int16_t hm = h1 < h2 ? h1 : h2; hm = hm < h3 ? hm : h3;
        if (i0 < GO2_NUM_HEIGHT) S.heights[i] = (float)hm * C->vscale;
      }
    }
    if (lane == 0) {
      S.ep_len += 1;
      S.resamp_step -= 1.0f;
      if (C->turn_over) { float* tt = GO2_EXT_PTR(float*, C, ext_turn_over_timer) + e; *tt = fmaxf(*tt - C->dt, 0.0f); }   // legged_robot.py:114-115
      V3 blv = quat_rotate_inverse(S.root + 3, ld3(S.root + 7)), bav = quat_rotate_inverse(S.root + 3, ld3(S.root + 10));
      V3 pg = quat_rotate_inverse(S.root + 3, mk(0.0f, 0.0f, -1.0f));
      st3(S.blv, blv); st3(S.bav, bav); st3(S.pg, pg);
      float dx = S.root[0] - S.env_origin[0], dy = S.root[1] - S.env_origin[1];
      S.max_move = fmaxf(S.max_move, sqrtf(dx * dx + dy * dy));
      if (S.resamp_step <= 0.0f && S.ep_len < C->max_episode_length - 1) resample_commands(S, X, e, ST_CMD_CB);
      if (C->heading_command && !S.stop_heading) heading_to_yaw(S);
    }
  } GO2_SYNC_WARP(); GO2_TICK();
  GO2_WIDE_COLD {
    {
      float sh = 0;
      GO2_STRIDED(i, GO2_NUM_HEIGHT) sh += S.heights[i] * CT->base_height_mask[i];
      S.part[lane] = sh;
    }
    if (lane < GO2_NUM_DOF) {  // per-joint reward terms
      float q = S.q[lane], qd = S.qd[lane], tq = S.tq[lane], a = S.act[lane], la = S.lact[lane], lla = S.llact[lane];
      float acc = (S.lqd[lane] - qd) / C->dt;
      S.jterm[0][lane] = acc * acc;
      S.jterm[1][lane] = fabsf(tq * qd);
      S.jterm[2][lane] = tq * tq;
      S.jterm[3][lane] = (la - a) * (la - a);
      float sm = a - 2 * la + lla;
      S.jterm[4][lane] = sm * sm;
      S.jterm[5][lane] = -fminf(q - CT->soft_dof_limit_lo[lane], 0.0f) + fmaxf(q - CT->soft_dof_limit_hi[lane], 0.0f);
      S.jterm[6][lane] = (lane % 3 == 0) ? fabsf(q - L.ddp) : 0.0f;
      S.llact[lane] = la;  // legged_robot.py:1378
    }
    if (lane >= 16 && lane < 24) {  // collision flags: thigh, calf of each leg (reported bodies 4+4l, 5+4l)
      const int l = (lane - 16) / 2, k = 1 + (lane - 16) % 2;
      const float* f = S.cf[3 + 4 * l + k];
      S.coll[lane - 16] = (sqrtf(f[0] * f[0] + f[1] * f[1] + f[2] * f[2]) > 0.1f) ? 1.0f : 0.0f;
    }
  } GO2_SYNC_WARP(); GO2_TICK();
  GO2_WIDE_COLD {
    if (lane == 0) {
      float sh = 0;
      for (int k = 0; k < 32; ++k) sh += S.part[k];
      S.base_height = S.root[2] - sh / C->num_base_height_points;
      const float* f = S.cf[0];
      int term = !C->turn_over && sqrtf(f[0] * f[0] + f[1] * f[1] + f[2] * f[2]) > 1.0f;     // no contact termination with turn_over (legged_robot.py:174-175)
      S.tout = S.ep_len > C->max_episode_length;
      S.reset = term || S.tout;
      S.reset |= S.bad;
    }
  } GO2_SYNC_WARP(); GO2_TICK();
  GO2_WIDE_COLD {
    if (lane < 4) {  // feet_regulation per foot, legged_robot.py:1404-1414
      const float* fp = S.feet[lane];
      float f2b = (fp[0] - S.root[0]) * S.pg[0] + (fp[1] - S.root[1]) * S.pg[1] + (fp[2] - S.root[2]) * S.pg[2];
      float fh = fmaxf(S.base_height - f2b, 0.0f);
      S.fterm[lane] = (fp[3] * fp[3] + fp[4] * fp[4]) * expf(-fh / (0.025f * C->base_height_target));
    }
  } GO2_SYNC_WARP(); GO2_TICK();
  GO2_WIDE_COLD {
    if (lane == 0) {
      float tv[GO2_NUM_REW];
      float ds_on = (C->terrain_curriculum && C->dynamic_sigma && C->mesh_type != 0) ? 1.0f : 0.0f;
      float sig[3];
      for (int a = 0; a < 3; ++a) {  // _get_dynamic_sigma, legged_robot.py:1300-1320
        float d = C->tracking_sigma, sgm = d;
        if (ds_on != 0.0f) {
          float tvel = fabsf(S.cmd[a]), vmin = a < 2 ? C->ds_min_lin : C->ds_min_ang, vmax = a < 2 ? C->ds_max_lin : C->ds_max_ang;
          float target = C->ds_max_sigma[S.tid];
          if (tvel >= vmin && tvel < vmax) sgm = d + (tvel - vmin) / (vmax - vmin) * (target - d);
          if (tvel >= vmax) sgm = target;
          float ls = fminf(expf(((float)S.level + 1.0f) / 10.0f) - 1.0f, 1.0f);
          sgm = d + ls * (sgm - d);
        }
        sig[a] = sgm;
      }
      float ex = (S.cmd[0] - S.blv[0]) * (S.cmd[0] - S.blv[0]), ey = (S.cmd[1] - S.blv[1]) * (S.cmd[1] - S.blv[1]);
      tv[GO2_REW_TRACKING_LIN_VEL] = expf(-(ex / sig[0] + ey / sig[1]));
      tv[GO2_REW_TRACKING_ANG_VEL] = expf(-((S.cmd[2] - S.bav[2]) * (S.cmd[2] - S.bav[2])) / sig[2]);
      tv[GO2_REW_LIN_VEL_Z] = S.blv[2] * S.blv[2];
      tv[GO2_REW_ANG_VEL_XY] = S.bav[0] * S.bav[0] + S.bav[1] * S.bav[1];
      float js[7];
      for (int t = 0; t < 7; ++t) { float a = 0; for (int j = 0; j < GO2_NUM_DOF; ++j) a += S.jterm[t][j]; js[t] = a; }
      tv[GO2_REW_DOF_ACC] = js[0]; tv[GO2_REW_DOF_POWER] = js[1]; tv[GO2_REW_TORQUES] = js[2]; tv[GO2_REW_ACTION_RATE] = js[3];
      tv[GO2_REW_ACTION_SMOOTHNESS] = js[4]; tv[GO2_REW_DOF_POS_LIMITS] = js[5]; tv[GO2_REW_HIP_TO_DEFAULT] = js[6];
      tv[GO2_REW_CORRECT_BASE_HEIGHT] = (S.base_height - C->base_height_target) * (S.base_height - C->base_height_target);
      float cs = 0; for (int k = 0; k < 8; ++k) cs += S.coll[k];
      tv[GO2_REW_COLLISION] = cs;
      tv[GO2_REW_FEET_REGULATION] = S.fterm[0] + S.fterm[1] + S.fterm[2] + S.fterm[3];
      float rew = 0;
      bool need_to = false;     // turn_over: while |roll| exceeds the threshold every term uses its turn_over scale (legged_robot.py:257-265; get_euler_xyz, isaacgym_utils.py:11-17)
      if (C->turn_over) {
        const float qx = S.root[3], qy = S.root[4], qz = S.root[5], qw = S.root[6];
        need_to = fabsf(atan2f(2.0f * (qw * qx + qy * qz), qw * qw - qx * qx - qy * qy + qz * qz)) > C->turn_over_roll_threshold;
      }
      for (int k = 0; k < GO2_NUM_REW; ++k) {
        float rk = tv[k] * (need_to ? C->to_scales[k] : C->reward_scales[k]) * sp->reward_curriculum[k];
        rew += rk;
        S.termv[k] = rk;
      }
      float term_rew = 0.0f;
      if (C->num_xrew > 0) extra_rewards(S, X, e, need_to, rew, term_rew);
      if (C->only_positive_rewards) rew = fmaxf(rew, 0.0f);     // the episode sums keep the unclipped terms (legged_robot.py:263-267)
      S.rew = rew + term_rew;                                   // termination reward after the clip (legged_robot.py:268-272)
    }
  } GO2_SYNC_WARP(); GO2_TICK();
  GO2_WIDE_COLD {
    if (lane < GO2_NUM_REW) B->episode_sums[(size_t)e * GO2_NUM_REW + lane] += S.termv[lane];
    if (lane == 31) B->rew_buf[e] = S.rew;
  } GO2_SYNC_WARP(); GO2_TICK();
  if (GO2_ANY_RESET(SM)) reset_phases<T>(GO2_LANE_PASS, SM, X, false);
  GO2_TICK();   // warp-uniform (WIDE phases only); items are predicated by their env
  GO2_WIDE_COLD {
    if (lane == 0) {
      if (C->push_robots && (S.ep_len % C->push_interval == 0)) {  // _push_robots, legged_robot.py:709-724
        const uint32_t ge = (uint32_t)(C->env_offset + e);
        U4 p0 = philox(ge, sp->common_step_counter, ST_PUSH, 0, C->seed_lo, C->seed_hi);
        U4 p1 = philox(ge, sp->common_step_counter, ST_PUSH, 1, C->seed_lo, C->seed_hi);
        float mv = C->max_push_vel_xy, ma = C->max_push_ang_vel;
        S.root[7] = affine(2 * mv, u01(p0.x), -mv); S.root[8] = affine(2 * mv, u01(p0.y), -mv);
        S.root[10] = affine(2 * ma, u01(p0.z), -ma); S.root[11] = affine(2 * ma, u01(p0.w), -ma);
        S.root[12] = affine(2 * ma, u01(p1.x), -ma);
      }
      // cross-env logging sums (extras["episode"], legged_robot.py:229-235): levels AFTER the curriculum update
      GO2_ATOMIC_ADD(B->ep_accum + GO2_NUM_REW, (float)S.level);
      GO2_ATOMIC_ADD(B->ep_accum + GO2_NUM_REW + 1 + S.tid, (float)S.level);
    }
  } GO2_SYNC_WARP(); GO2_TICK();
  // ---- compute_observations (go2_env.py:23-53) + clip (legged_robot.py:96-99).  The 76 proprioceptive columns are first written to
  // shared memory by the lanes that own their sources (no divergent 12-way branch per column), then rows leave coalesced.
  GO2_WIDE_COLD {
    if (lane < GO2_NUM_DOF) {
      S.obsrow[12 + lane] = (S.q[lane] - L.ddp) * C->obs_scale_dof_pos;
      S.obsrow[24 + lane] = S.qd[lane] * C->obs_scale_dof_vel;
      S.obsrow[36 + lane] = S.act[lane];
      S.obsrow[52 + lane] = S.tq[lane] / L.eff;
      S.obsrow[64 + lane] = (S.lqd[lane] - S.qd[lane]) / C->dt * 1e-4f;
    } else if (lane < 15) {
      const int k = lane - 12;
      S.obsrow[k] = S.blv[k] * C->obs_scale_lin_vel; S.obsrow[3 + k] = S.bav[k] * C->obs_scale_ang_vel; S.obsrow[6 + k] = S.pg[k];
    } else if (lane < 18) {
      const int k = lane - 15;
      S.obsrow[9 + k] = S.cmd[k] * (k < 2 ? C->obs_scale_lin_vel : C->obs_scale_ang_vel);
    } else if (lane < 22) {
      const float* f = S.cf[6 + 4 * (lane - 18)];
      S.obsrow[48 + lane - 18] = sqrtf(f[0] * f[0] + f[1] * f[1] + f[2] * f[2]) * 1e-3f;
    }
  } GO2_SYNC_WARP(); GO2_TICK();
  GO2_WIDE_COLD {
#pragma unroll
    for (int k = 0; k < (GO2_NUM_PRIV + 31) / 32; ++k) {
      const int i = lane + 32 * k;
      if (i < GO2_NUM_PRIV) {
        float x = (i < 76) ? S.obsrow[i] : fminf(fmaxf(S.root[2] - 0.5f - S.heights[i < 76 ? 0 : i - 76], -1.0f), 1.0f) * C->obs_scale_height;
        B->privileged_obs_buf[(size_t)e * GO2_NUM_PRIV + i] = fminf(fmaxf(x, -C->clip_obs), C->clip_obs);
        if (k < 2 && i >= 3 && i < 48) {
          const int o = i - 3;
          if (C->add_noise) {
            const uint32_t ge = (uint32_t)(C->env_offset + e);
            U4 r = philox(ge, sp->common_step_counter, ST_NOISE, (uint32_t)(o / 4), C->seed_lo, C->seed_hi);
            x = GO2_FADD(x, GO2_FMUL(GO2_FADD(GO2_FMUL(2.0f, u01(pick(r, o % 4))), -1.0f), CT->noise_scale_vec[o]));
          }
          B->obs_buf[(size_t)e * GO2_NUM_OBS + o] = fminf(fmaxf(x, -C->clip_obs), C->clip_obs);
        }
      }
    }
    GO2_STRIDED(i, GO2_NUM_HEIGHT) B->measured_heights[(size_t)e * GO2_NUM_HEIGHT + i] = S.heights[i];
    if (lane < 3) { B->base_lin_vel[(size_t)e * 3 + lane] = S.blv[lane]; B->base_ang_vel[(size_t)e * 3 + lane] = S.bav[lane]; B->projected_gravity[(size_t)e * 3 + lane] = S.pg[lane]; }
  } GO2_SYNC_WARP(); GO2_TICK();
  GO2_WIDE_COLD {
    if (lane < GO2_NUM_DOF) { S.lact[lane] = S.act[lane]; S.lqd[lane] = S.qd[lane]; }
  } GO2_SYNC_WARP(); GO2_TICK();
  store_state<T>(GO2_LANE_PASS, SM, X); GO2_TICK();
}

// reset_idx(all envs) at construction (base_task.py:82-86; the zero-action step that follows is issued by the caller)
template <class T>
GO2_HD void reset_env_initial(GO2_LANE_ARGS, typename T::Smem* SM, const StepCtx& X) {
  load_env<T>(GO2_LANE_PASS, SM, X);
  GO2_WIDE {
    if (lane == 0) { S.tout = 0; }
    if (lane < 4) for (int k = 0; k < 6; ++k) S.feet[lane][k] = 0;
  } GO2_SYNC_WARP();
  reset_phases<T>(GO2_LANE_PASS, SM, X, true);
  feet_kinematics<T>(GO2_LANE_PASS, SM, X);
  store_state<T>(GO2_LANE_PASS, SM, X);
}

// n physics substeps with given joint torques (dynamics parity in isolation)
template <class T>
GO2_HD void substeps_env(GO2_LANE_ARGS, typename T::Smem* SM, const StepCtx& X, const float* tau_in, int n) {
  const Go2Model* M = X.mdl;
  load_env<T>(GO2_LANE_PASS, SM, X);
  for (int s = 0; s < n; ++s) {
    GO2_WIDE {
      if (lane < GO2_NUM_DOF) { float lim = M->effort[lane]; S.tau[lane] = fminf(fmaxf(tau_in[(size_t)e * GO2_NUM_DOF + lane], -lim), lim); }
    } GO2_SYNC_WARP();
    physics_substep<T>(GO2_LANE_PASS, SM, X, s == n - 1);
  }
  feet_kinematics<T>(GO2_LANE_PASS, SM, X);
  GO2_WIDE {
    if (lane == 0) { S.reset = X.buf->reset_buf[e]; S.tout = X.buf->time_out_buf[e]; }
  } GO2_SYNC_WARP();
  store_state<T>(GO2_LANE_PASS, SM, X);
}

}  // namespace go2

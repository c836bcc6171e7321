// go2_oracle.cpp — CPU ORACLE for the Go2 environment step.  TEST INFRASTRUCTURE ONLY.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this
// library; the product path (go2_rl_gym_b200/) never does and fails loudly without its CUDA extension.
//
// What it restates, scalar and one env at a time, with dense 6x6 spatial algebra (deliberately NOT the
// block-structured, warp-cooperative formulation of the CUDA kernel, so the two are independent):
//   * LeggedRobot.step / post_physics_step and everything they call
//     (/root/reference/legged_gym/envs/base/legged_robot.py:60-142, :170-245, :247-274, :404-421, :423-592,
//      :594-618, :620-724, :1143-1169, :1188-1224, rewards :1228-1414; go2_env.py:9-60;
//      utils/isaacgym_utils.py:32-55; utils/math.py:8-12).  These parts are PINNED against the reference's own
//      Python imported through an isaacgym stub (tests/golden/make_golden_env.py -> tests/golden/env_*.npz).
//   * gym.simulate (legged_robot.py:83): PhysX is closed source and absent -> PARITY UNPINNED for the physics.
//     The oracle follows the written physics spec of DESIGN.md section 3 instead (floating-base Featherstone ABA,
//     semi-implicit Euler at 5 ms, sphere-sample colliders vs bilinear heightfield, velocity-level contact and
//     joint-limit impulses solved by mass-split Jacobi sweeps with exact tree propagation).
//   * Random draws: Philox4x32-10 keyed by (seed; global env, step, stream, block) — the reference's torch
//     global generator with data-dependent shapes cannot be replayed.
//
// Build: make -C oracle   (float build libgo2oracle.so, double build libgo2oracle_f64.so)
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <algorithm>
#include <vector>

#include "../include/go2_b200.h"

#ifdef ORACLE_DOUBLE
typedef double real;
#else
typedef float real;
#endif

namespace {

// ------------------------------------------------------------------------------------------------ Philox
struct U4 { uint32_t x, y, z, w; };
inline U4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
  for (int r = 0; r < 10; ++r) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  return U4{c0, c1, c2, c3};
}
inline float u01(uint32_t x) { return (float)(x >> 8) * (1.0f / 16777216.0f); }
enum Stream { ST_DELAY = 0, ST_NOISE = 1, ST_PUSH = 2, ST_RESET_DR = 3, ST_RESET_STATE = 4, ST_CMD_CB = 5, ST_CMD_RESET = 6 };

// ------------------------------------------------------------------------------------------------ small algebra
struct V3 { real x, y, z; };
inline V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 operator*(real s, V3 a) { return {s * a.x, s * a.y, s * a.z}; }
inline real dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline V3 cross(V3 a, V3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
inline real norm(V3 a) { return std::sqrt(dot(a, a)); }
struct M3 { real m[3][3]; };
inline V3 mul(const M3& A, V3 v) {
  return {A.m[0][0] * v.x + A.m[0][1] * v.y + A.m[0][2] * v.z, A.m[1][0] * v.x + A.m[1][1] * v.y + A.m[1][2] * v.z,
          A.m[2][0] * v.x + A.m[2][1] * v.y + A.m[2][2] * v.z};
}
inline V3 mulT(const M3& A, V3 v) {
  return {A.m[0][0] * v.x + A.m[1][0] * v.y + A.m[2][0] * v.z, A.m[0][1] * v.x + A.m[1][1] * v.y + A.m[2][1] * v.z,
          A.m[0][2] * v.x + A.m[1][2] * v.y + A.m[2][2] * v.z};
}
inline M3 mul(const M3& A, const M3& B) {
  M3 C;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) C.m[i][j] = A.m[i][0] * B.m[0][j] + A.m[i][1] * B.m[1][j] + A.m[i][2] * B.m[2][j];
  return C;
}
inline M3 transpose(const M3& A) {
  M3 C;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) C.m[i][j] = A.m[j][i];
  return C;
}
inline M3 skew(V3 r) { return M3{{{0, -r.z, r.y}, {r.z, 0, -r.x}, {-r.y, r.x, 0}}}; }
inline M3 eye3() { return M3{{{1, 0, 0}, {0, 1, 0}, {0, 0, 1}}}; }
inline M3 rot_axis(int axis, real q) {  // rotation matrix of angle q about a coordinate axis (child -> parent coords)
  real c = std::cos(q), s = std::sin(q);
  if (axis == 0) return M3{{{1, 0, 0}, {0, c, -s}, {0, s, c}}};
  if (axis == 1) return M3{{{c, 0, s}, {0, 1, 0}, {-s, 0, c}}};
  return M3{{{c, -s, 0}, {s, c, 0}, {0, 0, 1}}};
}
inline M3 quat_to_mat(const real* q) {  // xyzw, body -> world
  real x = q[0], y = q[1], z = q[2], w = q[3];
  return M3{{{1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)},
             {2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)},
             {2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)}}};
}

struct V6 { real v[6]; };
struct M6 { real m[6][6]; };
inline V6 zero6() { V6 a; for (int i = 0; i < 6; ++i) a.v[i] = 0; return a; }
inline M6 zero66() { M6 a; std::memset(&a, 0, sizeof a); return a; }
inline V6 mul(const M6& A, const V6& x) {
  V6 y;
  for (int i = 0; i < 6; ++i) { real s = 0; for (int j = 0; j < 6; ++j) s += A.m[i][j] * x.v[j]; y.v[i] = s; }
  return y;
}
inline V6 mulT(const M6& A, const V6& x) {
  V6 y;
  for (int i = 0; i < 6; ++i) { real s = 0; for (int j = 0; j < 6; ++j) s += A.m[j][i] * x.v[j]; y.v[i] = s; }
  return y;
}
inline M6 mul(const M6& A, const M6& B) {
  M6 C;
  for (int i = 0; i < 6; ++i)
    for (int j = 0; j < 6; ++j) { real s = 0; for (int k = 0; k < 6; ++k) s += A.m[i][k] * B.m[k][j]; C.m[i][j] = s; }
  return C;
}
inline M6 transpose(const M6& A) { M6 C; for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) C.m[i][j] = A.m[j][i]; return C; }
inline V6 add(const V6& a, const V6& b) { V6 c; for (int i = 0; i < 6; ++i) c.v[i] = a.v[i] + b.v[i]; return c; }
inline real dot(const V6& a, const V6& b) { real s = 0; for (int i = 0; i < 6; ++i) s += a.v[i] * b.v[i]; return s; }
inline V3 ang(const V6& a) { return {a.v[0], a.v[1], a.v[2]}; }
inline V3 lin(const V6& a) { return {a.v[3], a.v[4], a.v[5]}; }
inline V6 mk6(V3 a, V3 l) { return V6{{a.x, a.y, a.z, l.x, l.y, l.z}}; }
// Plücker motion transform parent -> child: X = [E 0; -E rx, E], E = parent->child rotation, r = child origin in parent
inline M6 plucker(const M3& E, V3 r) {
  M6 X = zero66();
  M3 Erx = mul(E, skew(r));
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) { X.m[i][j] = E.m[i][j]; X.m[i + 3][j + 3] = E.m[i][j]; X.m[i + 3][j] = -Erx.m[i][j]; }
  return X;
}
inline V6 crm(const V6& v, const V6& m) {  // v x m (motion)
  V3 w = ang(v), vl = lin(v), mw = ang(m), ml = lin(m);
  return mk6(cross(w, mw), cross(w, ml) + cross(vl, mw));
}
inline V6 crf(const V6& v, const V6& f) {  // v x* f (force)
  V3 w = ang(v), vl = lin(v), fn = ang(f), fl = lin(f);
  return mk6(cross(w, fn) + cross(vl, fl), cross(w, fl));
}
inline M6 spatial_inertia(const float* rec) {  // rec: mass, com xyz, Ixx Iyy Izz Ixy Ixz Iyz about COM
  real m = rec[0];
  V3 c{(real)rec[1], (real)rec[2], (real)rec[3]};
  M3 Ic{{{(real)rec[4], (real)rec[7], (real)rec[8]}, {(real)rec[7], (real)rec[5], (real)rec[9]}, {(real)rec[8], (real)rec[9], (real)rec[6]}}};
  M3 cx = skew(c), cxT = transpose(cx), cc = mul(cx, cxT);
  M6 I = zero66();
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      I.m[i][j] = Ic.m[i][j] + m * cc.m[i][j];
      I.m[i][j + 3] = m * cx.m[i][j];
      I.m[i + 3][j] = m * cxT.m[i][j];
      I.m[i + 3][j + 3] = (i == j) ? m : 0;
    }
  return I;
}
// solve A x = b for SPD 6x6 (Cholesky); also returns inverse if inv != nullptr
inline void chol6(const M6& A, real L[6][6]) {
  for (int i = 0; i < 6; ++i)
    for (int j = 0; j <= i; ++j) {
      real s = A.m[i][j];
      for (int k = 0; k < j; ++k) s -= L[i][k] * L[j][k];
      L[i][j] = (i == j) ? std::sqrt(s) : s / L[j][j];
    }
}
inline V6 chol_solve(const real L[6][6], const V6& b) {
  V6 y, x;
  for (int i = 0; i < 6; ++i) { real s = b.v[i]; for (int k = 0; k < i; ++k) s -= L[i][k] * y.v[k]; y.v[i] = s / L[i][i]; }
  for (int i = 5; i >= 0; --i) { real s = y.v[i]; for (int k = i + 1; k < 6; ++k) s -= L[k][i] * x.v[k]; x.v[i] = s / L[i][i]; }
  return x;
}

// ------------------------------------------------------------------------------------------------ terrain
struct Terrain {
  const Go2EnvConfig* cfg;
  const int16_t* hs;
  // bilinear height and gradient of the heightfield at world (x, y); plane -> 0
  void query(real x, real y, real& h, real& dhdx, real& dhdy) const {
    if (cfg->mesh_type == 0) { h = 0; dhdx = 0; dhdy = 0; return; }
    real gx = (x + (real)cfg->border) / (real)cfg->hscale, gy = (y + (real)cfg->border) / (real)cfg->hscale;
    int ix = (int)std::floor(gx), iy = (int)std::floor(gy);
    ix = std::min(std::max(ix, 0), cfg->hf_rows - 2);
    iy = std::min(std::max(iy, 0), cfg->hf_cols - 2);
    real fx = std::min(std::max(gx - (real)ix, (real)0), (real)1), fy = std::min(std::max(gy - (real)iy, (real)0), (real)1);
    real h00 = hs[(size_t)ix * cfg->hf_cols + iy], h10 = hs[(size_t)(ix + 1) * cfg->hf_cols + iy];
    real h01 = hs[(size_t)ix * cfg->hf_cols + iy + 1], h11 = hs[(size_t)(ix + 1) * cfg->hf_cols + iy + 1];
    real vs = cfg->vscale, k = vs / (real)cfg->hscale;
    h = vs * ((1 - fx) * (1 - fy) * h00 + fx * (1 - fy) * h10 + (1 - fx) * fy * h01 + fx * fy * h11);
    dhdx = k * ((1 - fy) * (h10 - h00) + fy * (h11 - h01));
    dhdy = k * ((1 - fx) * (h01 - h00) + fx * (h11 - h10));
  }
};

// ------------------------------------------------------------------------------------------------ physics substep
struct Kin {            // per dynamic body kinematics of one configuration
  M3 Rw[GO2_NUM_DYN];   // body -> world
  V3 pw[GO2_NUM_DYN];   // origin in world
  M6 X[GO2_NUM_DYN];    // parent -> body Plücker transform (X[0] unused)
};
inline int parent_of(const Go2Model& M, int body) { return body == 0 ? -1 : ((body - 1) % 3 == 0 ? 0 : body - 1); }

static void kinematics(const Go2Model& M, const real* pos, const real* quat, const real* q, Kin& K) {
  K.Rw[0] = quat_to_mat(quat);
  K.pw[0] = {pos[0], pos[1], pos[2]};
  for (int b = 1; b < GO2_NUM_DYN; ++b) {
    int j = b - 1, p = parent_of(M, b);
    V3 r{(real)M.joint_origin[j][0], (real)M.joint_origin[j][1], (real)M.joint_origin[j][2]};
    M3 Rpc = rot_axis(M.joint_axis[j], q[j]);
    K.X[b] = plucker(transpose(Rpc), r);
    K.Rw[b] = mul(K.Rw[p], Rpc);
    K.pw[b] = K.pw[p] + mul(K.Rw[p], r);
  }
}

struct SubstepOut { real contact_force[GO2_NUM_REPORT][3]; };

// One 5 ms step of the articulated body. pos/quat/lin/angvel are world-frame root state, q/qd joints, tau torques.
static void physics_substep(const Go2EnvConfig& C, const Go2Model& M, const Terrain& T, const float* inertia,
                            real mu_robot, real rest_robot, real* pos, real* quat, real* linw, real* angw, real* q,
                            real* qd, const real* tau, SubstepOut& out) {
  const real dt = C.sim_dt;
  Kin K;
  kinematics(M, pos, quat, q, K);
  // --- ABA pass 1: velocities and bias terms (body coordinates)
  V6 v[GO2_NUM_DYN], c[GO2_NUM_DYN], pA[GO2_NUM_DYN];
  M6 IA[GO2_NUM_DYN];
  V3 wb = mulT(K.Rw[0], V3{angw[0], angw[1], angw[2]}), vb = mulT(K.Rw[0], V3{linw[0], linw[1], linw[2]});
  v[0] = mk6(wb, vb);
  c[0] = zero6();
  for (int b = 0; b < GO2_NUM_DYN; ++b) {
    if (b > 0) {
      int p = parent_of(M, b);
      V6 vj = zero6();
      vj.v[M.joint_axis[b - 1]] = qd[b - 1];
      v[b] = add(mul(K.X[b], v[p]), vj);
      c[b] = crm(v[b], vj);
    }
    IA[b] = spatial_inertia(inertia + b * GO2_INERTIA_STRIDE);
    pA[b] = crf(v[b], mul(IA[b], v[b]));
  }
  // --- pass 2: articulated inertias, leaves to root
  V6 U[GO2_NUM_DYN];
  real D[GO2_NUM_DYN], u[GO2_NUM_DYN];
  for (int b = GO2_NUM_DYN - 1; b >= 1; --b) {
    int k = M.joint_axis[b - 1], p = parent_of(M, b);
    for (int i = 0; i < 6; ++i) U[b].v[i] = IA[b].m[i][k];
    D[b] = U[b].v[k];
    u[b] = tau[b - 1] - pA[b].v[k];
    M6 Ia = IA[b];
    for (int i = 0; i < 6; ++i)
      for (int j = 0; j < 6; ++j) Ia.m[i][j] -= U[b].v[i] * U[b].v[j] / D[b];
    V6 pa = add(pA[b], mul(Ia, c[b]));
    for (int i = 0; i < 6; ++i) pa.v[i] += U[b].v[i] * u[b] / D[b];
    M6 Xt = transpose(K.X[b]);
    M6 add_I = mul(Xt, mul(Ia, K.X[b]));
    V6 add_p = mul(Xt, pa);
    for (int i = 0; i < 6; ++i) {
      pA[p].v[i] += add_p.v[i];
      for (int j = 0; j < 6; ++j) IA[p].m[i][j] += add_I.m[i][j];
    }
  }
  real L0[6][6];
  chol6(IA[0], L0);
  V6 a[GO2_NUM_DYN];
  {
    V6 rhs = pA[0];
    for (int i = 0; i < 6; ++i) rhs.v[i] = -rhs.v[i];
    a[0] = chol_solve(L0, rhs);
  }
  // --- pass 3: accelerations (gravity-free frame), then add gravity to the base
  real qdd[GO2_NUM_DOF];
  for (int b = 1; b < GO2_NUM_DYN; ++b) {
    int k = M.joint_axis[b - 1], p = parent_of(M, b);
    V6 ap = add(mul(K.X[b], a[p]), c[b]);
    qdd[b - 1] = (u[b] - dot(U[b], ap)) / D[b];
    a[b] = ap;
    a[b].v[k] += qdd[b - 1];
  }
  V3 gb = mulT(K.Rw[0], V3{0, 0, (real)C.gravity_z});
  // --- unconstrained velocity update (semi-implicit Euler)
  V6 v0m = v[0];
  for (int i = 0; i < 6; ++i) v0m.v[i] += dt * a[0].v[i];
  v0m.v[3] += dt * gb.x; v0m.v[4] += dt * gb.y; v0m.v[5] += dt * gb.z;
  {  // components stay in the frame of the START of the step (K.Rw[0]): classical accel = spatial accel + w x v
    V3 wxv = cross(wb, vb);
    v0m.v[3] += dt * wxv.x; v0m.v[4] += dt * wxv.y; v0m.v[5] += dt * wxv.z;
  }
  real qdm[GO2_NUM_DOF];
  for (int j = 0; j < GO2_NUM_DOF; ++j) qdm[j] = qd[j] + dt * qdd[j];
  V6 vm[GO2_NUM_DYN];
  vm[0] = v0m;
  for (int b = 1; b < GO2_NUM_DYN; ++b) {
    vm[b] = mul(K.X[b], vm[parent_of(M, b)]);
    vm[b].v[M.joint_axis[b - 1]] += qdm[b - 1];
  }
  // --- operational-space inverse inertia (mobility) of every body, root to leaves
  M6 Lam[GO2_NUM_DYN];
  for (int col = 0; col < 6; ++col) {
    V6 e = zero6(); e.v[col] = 1;
    V6 x = chol_solve(L0, e);
    for (int i = 0; i < 6; ++i) Lam[0].m[i][col] = x.v[i];
  }
  for (int b = 1; b < GO2_NUM_DYN; ++b) {
    int k = M.joint_axis[b - 1], p = parent_of(M, b);
    M6 LtX = K.X[b];                       // L^T X, L^T = 1 - S U^T / D
    V6 UtX = mulT(K.X[b], U[b]);           // (U^T X)^T
    for (int j = 0; j < 6; ++j) LtX.m[k][j] -= UtX.v[j] / D[b];
    Lam[b] = mul(LtX, mul(Lam[p], transpose(LtX)));
    Lam[b].m[k][k] += 1 / D[b];
  }
  // --- contact candidates
  struct Contact { int body, rep; V3 r, n; real vt, mu; V3 p; M3 Winv; bool active; };
  Contact ct[GO2_NUM_COL];
  int group_count[5] = {0, 0, 0, 0, 0};  // base, leg0..3
  const real mu = (C.terrain_friction + mu_robot) / 2, rest = (C.terrain_restitution + rest_robot) / 2;
  for (int ci = 0; ci < GO2_NUM_COL; ++ci) {
    Contact& k = ct[ci];
    k.body = M.col_dyn[ci]; k.rep = M.col_report[ci];
    k.r = {(real)M.col_pos[ci][0], (real)M.col_pos[ci][1], (real)M.col_pos[ci][2]};
    k.p = {0, 0, 0};
    V3 cw = K.pw[k.body] + mul(K.Rw[k.body], k.r);
    real h, dhx, dhy;
    T.query(cw.x, cw.y, h, dhx, dhy);
    real inv = 1 / std::sqrt(dhx * dhx + dhy * dhy + 1);
    k.n = {-dhx * inv, -dhy * inv, inv};
    real gap = (cw.z - h) * k.n.z - (real)M.col_radius[ci];
    k.active = gap < (real)C.contact_offset;
    if (!k.active) continue;
    group_count[k.body == 0 ? 0 : 1 + (k.body - 1) / 3]++;
    // point Jacobian (world velocity of the point = Rw [-rx 1] v_body)
    M3 rx = skew(k.r);
    M6 Lb = Lam[k.body];
    // W = J Lam J^T with J = Rw [ -rx , 1 ]   (3x6)
    real Jm[3][6];
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) {
        real s0 = 0;
        for (int t = 0; t < 3; ++t) s0 += K.Rw[k.body].m[i][t] * (-rx.m[t][j]);
        Jm[i][j] = s0;
        Jm[i][j + 3] = K.Rw[k.body].m[i][j];
      }
    real W[3][3];
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) {
        real s0 = 0;
        for (int a1 = 0; a1 < 6; ++a1)
          for (int b1 = 0; b1 < 6; ++b1) s0 += Jm[i][a1] * Lb.m[a1][b1] * Jm[j][b1];
        W[i][j] = s0;
      }
    {  // closed-form inverse of the symmetric positive definite 3x3 point mobility
      real a = W[0][0], b = W[0][1], c2 = W[0][2], d = W[1][1], e2 = W[1][2], f = W[2][2];
      real c00 = d * f - e2 * e2, c01 = c2 * e2 - b * f, c02 = b * e2 - c2 * d;
      real det = a * c00 + b * c01 + c2 * c02, id = 1 / det;
      k.Winv = M3{{{c00 * id, c01 * id, c02 * id}, {c01 * id, (a * f - c2 * c2) * id, (b * c2 - a * e2) * id},
                   {c02 * id, (b * c2 - a * e2) * id, (a * d - b * b) * id}}};
    }
    k.mu = mu;
    // velocity target along the normal
    // restitution looks at the approach speed at the START of the step (before this step's forces act)
    V3 vp = mul(K.Rw[k.body], lin(v[k.body]) + cross(ang(v[k.body]), k.r));
    real vn = dot(vp, k.n);
    real vt = (gap >= 0) ? -gap / dt : std::min(std::max(-gap - (real)C.penetration_slop, (real)0) * (real)C.erp / dt, (real)C.max_depen_vel);
    if (vn < -(real)C.bounce_threshold) vt = std::max(vt, -rest * vn);
    k.vt = vt;
  }
  // --- joint limit rows
  real lam_lo[GO2_NUM_DOF], lam_hi[GO2_NUM_DOF], tgt_lo[GO2_NUM_DOF], tgt_hi[GO2_NUM_DOF];
  for (int j = 0; j < GO2_NUM_DOF; ++j) {
    lam_lo[j] = lam_hi[j] = 0;
    real glo = q[j] - (real)M.q_lower[j], ghi = (real)M.q_upper[j] - q[j];
    tgt_lo[j] = (glo >= 0) ? -glo / dt : -glo * (real)C.limit_erp / dt;   // qd+ >= tgt_lo
    tgt_hi[j] = (ghi >= 0) ? ghi / dt : ghi * (real)C.limit_erp / dt;     // qd+ <= tgt_hi
  }
  // --- Jacobi sweeps with exact propagation through the tree
  V6 dv[GO2_NUM_DYN];
  real dqd[GO2_NUM_DOF];
  for (int b = 0; b < GO2_NUM_DYN; ++b) dv[b] = zero6();
  for (int j = 0; j < GO2_NUM_DOF; ++j) dqd[j] = 0;
  for (int it = 0; it < C.solver_iters; ++it) {
    for (int ci = 0; ci < GO2_NUM_COL; ++ci) {
      Contact& k = ct[ci];
      if (!k.active) continue;
      real s = (real)group_count[k.body == 0 ? 0 : 1 + (k.body - 1) / 3];
      V6 vb6 = add(vm[k.body], dv[k.body]);
      V3 vp = mul(K.Rw[k.body], lin(vb6) + cross(ang(vb6), k.r));
      // block solve toward (normal velocity = target, zero slip), split by the group's contact count, then
      // project the accumulated impulse onto the friction cone
      V3 err = vp - k.vt * k.n;
      V3 pc = k.p - ((real)C.contact_relax / s) * mul(k.Winv, err);
      real pn = std::max((real)0, dot(pc, k.n));
      V3 pt = pc - dot(pc, k.n) * k.n;
      real ptn = norm(pt), lim = k.mu * pn;
      if (ptn > lim) pt = (ptn > 0 ? lim / ptn : (real)0) * pt;
      k.p = pn * k.n + pt;
    }
    for (int j = 0; j < GO2_NUM_DOF; ++j) {
      real cur = qdm[j] + dqd[j], Dj = D[j + 1];
      if (C.limit_relax > 0) Dj = (real)C.limit_relax / Lam[j + 1].m[M.joint_axis[j]][M.joint_axis[j]];   // exact joint-space diagonal
      lam_lo[j] = std::max((real)0, lam_lo[j] + (tgt_lo[j] - cur) * Dj);
      lam_hi[j] = std::min((real)0, lam_hi[j] + (tgt_hi[j] - cur) * Dj);
    }
    // impulse propagation: p_i = -f_i, leaves to root, then root to leaves
    V6 pI[GO2_NUM_DYN];
    real uI[GO2_NUM_DYN];
    for (int b = 0; b < GO2_NUM_DYN; ++b) pI[b] = zero6();
    for (int ci = 0; ci < GO2_NUM_COL; ++ci) {
      const Contact& k = ct[ci];
      if (!k.active) continue;
      V3 fl = mulT(K.Rw[k.body], k.p);
      V3 fn = cross(k.r, fl);
      V6 f = mk6(fn, fl);
      for (int i = 0; i < 6; ++i) pI[k.body].v[i] -= f.v[i];
    }
    for (int b = GO2_NUM_DYN - 1; b >= 1; --b) {
      int kx = M.joint_axis[b - 1], p = parent_of(M, b);
      uI[b] = (lam_lo[b - 1] + lam_hi[b - 1]) - pI[b].v[kx];
      V6 pa = pI[b];
      for (int i = 0; i < 6; ++i) pa.v[i] += U[b].v[i] * uI[b] / D[b];
      V6 add_p = mulT(K.X[b], pa);
      for (int i = 0; i < 6; ++i) pI[p].v[i] += add_p.v[i];
    }
    {
      V6 rhs = pI[0];
      for (int i = 0; i < 6; ++i) rhs.v[i] = -rhs.v[i];
      dv[0] = chol_solve(L0, rhs);
    }
    for (int b = 1; b < GO2_NUM_DYN; ++b) {
      int kx = M.joint_axis[b - 1], p = parent_of(M, b);
      V6 dp = mul(K.X[b], dv[p]);
      dqd[b - 1] = (uI[b] - dot(U[b], dp)) / D[b];
      dv[b] = dp;
      dv[b].v[kx] += dqd[b - 1];
    }
  }
  // --- final velocities, joint velocity clamp, integrate
  V6 v0p = add(vm[0], dv[0]);
  for (int j = 0; j < GO2_NUM_DOF; ++j) {
    real x = qdm[j] + dqd[j], vl = M.vel_limit[j];
    qd[j] = std::min(std::max(x, -vl), vl);
    q[j] += dt * qd[j];
  }
  V3 lw = mul(K.Rw[0], lin(v0p)), aw = mul(K.Rw[0], ang(v0p));
  if (C.state_guard) {   // asset.max_linear_velocity / max_angular_velocity (legged_robot_config.py:131-132)
    const real nl = norm(lw), na = norm(aw);
    if (nl > (real)C.max_base_lin_vel) lw = ((real)C.max_base_lin_vel / nl) * lw;
    if (na > (real)C.max_base_ang_vel) aw = ((real)C.max_base_ang_vel / na) * aw;
  }
  linw[0] = lw.x; linw[1] = lw.y; linw[2] = lw.z;
  angw[0] = aw.x; angw[1] = aw.y; angw[2] = aw.z;
  pos[0] += dt * lw.x; pos[1] += dt * lw.y; pos[2] += dt * lw.z;
  {
    real th = norm(aw) * dt, hx, hy, hz, hw;  // dq = exp(aw dt / 2), q <- dq * q (world-frame angular velocity)
    if (th > (real)1e-8) {
      real s = std::sin(th / 2) / norm(aw);
      hx = aw.x * s; hy = aw.y * s; hz = aw.z * s; hw = std::cos(th / 2);
    } else {
      hx = aw.x * dt / 2; hy = aw.y * dt / 2; hz = aw.z * dt / 2; hw = 1;
    }
    real x = quat[0], y = quat[1], z = quat[2], w = quat[3];
    real nx = hw * x + hx * w + hy * z - hz * y;
    real ny = hw * y - hx * z + hy * w + hz * x;
    real nz = hw * z + hx * y - hy * x + hz * w;
    real nw = hw * w - hx * x - hy * y - hz * z;
    real nn = std::sqrt(nx * nx + ny * ny + nz * nz + nw * nw);
    quat[0] = nx / nn; quat[1] = ny / nn; quat[2] = nz / nn; quat[3] = nw / nn;
  }
  for (int b = 0; b < GO2_NUM_REPORT; ++b) out.contact_force[b][0] = out.contact_force[b][1] = out.contact_force[b][2] = 0;
  for (int ci = 0; ci < GO2_NUM_COL; ++ci) {
    const Contact& k = ct[ci];
    if (!k.active) continue;
    out.contact_force[k.rep][0] += k.p.x / dt;
    out.contact_force[k.rep][1] += k.p.y / dt;
    out.contact_force[k.rep][2] += k.p.z / dt;
  }
}

// ------------------------------------------------------------------------------------------------ helpers (torch_utils restated)
inline void quat_rotate_inverse(const real* q, const real* v, real* o) {
  real w = q[3];
  V3 qv{q[0], q[1], q[2]}, vv{v[0], v[1], v[2]};
  V3 a = (2 * w * w - 1) * vv, b = (2 * w) * cross(qv, vv), c = (2 * dot(qv, vv)) * qv;
  o[0] = a.x - b.x + c.x; o[1] = a.y - b.y + c.y; o[2] = a.z - b.z + c.z;
}

struct EnvView {  // pointers to the rows of one env
  const Go2EnvConfig* C; const Go2Model* M; const Go2EnvBuffers* B; int e;
};

static void resample_commands(const Go2EnvConfig& C, const Go2EnvBuffers& B, const Go2StepParams& sp, int e, int stream) {
  // legged_robot.py:423-592 for GO2Cfg (dynamic_resample_commands, no heading command)
  float* cmd = B.commands + (size_t)e * GO2_NUM_CMD;
  float* acc = B.commands_xy_accumulation + (size_t)e * 2;
  const float* rng = B.env_command_ranges + (size_t)e * 6;
  uint32_t ge = (uint32_t)(C.env_offset + e);
  U4 r0 = philox4x32_10(ge, sp.common_step_counter, (uint32_t)stream, 0, C.seed_lo, C.seed_hi);
  U4 r1 = philox4x32_10(ge, sp.common_step_counter, (uint32_t)stream, 1, C.seed_lo, C.seed_hi);
  float ep_len = (float)B.episode_length_buf[e];
  float max_len = (float)C.max_episode_length;
  float remaining = std::max(0.625f * C.terrain_length - std::sqrt(acc[0] * acc[0] + acc[1] * acc[1]) * C.resampling_time, 0.0f);
  B.commands_resampling_step[e] = C.resampling_time / C.dt;
  const bool heading = C.heading_command != 0;
  const float* hr = heading ? GO2_EXT_PTR(const float*, &C, ext_heading_ranges) + (size_t)e * 2 : nullptr;
  uint8_t* stop_heading = (heading || C.turn_over) ? GO2_EXT_PTR(uint8_t*, &C, ext_stop_heading) : nullptr;
  if (stop_heading) stop_heading[e] = 0;                          // legged_robot.py:431
  if (C.dynamic_resample_commands) {
    float vlow = std::max(remaining / ((max_len - ep_len + 1e-9f) * C.dt), 0.0f);
    for (int a = 0; a < 2; ++a) {  // sample_disjoint_intervals, isaacgym_utils.py:32-47
      float lo = rng[2 * a], hi = rng[2 * a + 1];
      float wneg = std::max(-vlow - lo, 0.0f), wpos = std::max(hi - vlow, 0.0f);
      float total = wneg + wpos + 1e-6f;
      float u = u01(a == 0 ? r0.x : r0.y) * total;
      cmd[a] = (u < wneg) ? lo + u : hi - wpos + (u - wneg);
    }
    if (heading) cmd[3] = (hr[1] - hr[0]) * u01(r0.z) + hr[0];          // the same draw feeds the heading target (:468-472)
    else cmd[2] = (rng[5] - rng[4]) * u01(r0.z) + rng[4];
  } else {
    cmd[0] = rng[0] + u01(r0.x) * (rng[1] - rng[0]);
    cmd[1] = rng[2] + u01(r0.y) * (rng[3] - rng[2]);
    if (heading) cmd[3] = hr[0] + u01(r0.z) * (hr[1] - hr[0]);
    else cmd[2] = rng[4] + u01(r0.z) * (rng[5] - rng[4]);
    float nrm = std::sqrt(cmd[0] * cmd[0] + cmd[1] * cmd[1]);
    if (!(nrm > 0.2f)) { cmd[0] = 0; cmd[1] = 0; }
  }
  float prob = u01(r0.w), min_p = 0, max_p = 0;
  if (C.limit_vel_prob > 0) {
    max_p += C.limit_vel_prob;
    bool lim = prob >= min_p && prob < max_p;
    if (lim) {
      bool change = true;
      if (C.limit_vel_invert_when_continuous && B.last_is_limit_vel[e]) {
        cmd[0] *= -1.0f; cmd[1] *= -1.0f; cmd[2] *= -1.0f;
        change = false;
      }
      if (change) {  // limit_vel_comb = product([-1,1],[-1,1],[-1,0,1]), legged_robot.py:827-831
        int idx = (int)(r1.x % 12u);
        int cx = idx / 6, cy = (idx / 3) % 2, cz = idx % 3;
        cmd[0] = cx == 0 ? rng[0] : rng[1];
        cmd[1] = cy == 0 ? rng[2] : rng[3];
        cmd[2] = cz == 0 ? rng[4] : (cz == 1 ? 0.0f : rng[5]);
      }
      if (heading && C.stop_heading_at_limit) stop_heading[e] = 1;   // :547-548
    }
    B.last_is_limit_vel[e] = lim ? 1 : 0;
    min_p += C.limit_vel_prob;
  }
  if (sp.zero_command_proba > 0) {
    max_p += sp.zero_command_proba;
    float next = max_len - ep_len - remaining / (0.8f * sp.max_lin_vel * C.dt + 1e-9f);
    next = std::min(std::max(next, 0.0f), C.resampling_time / C.dt);
    if (prob >= min_p && prob < max_p && next > 0) {
      cmd[0] = 0; cmd[1] = 0;
      B.commands_resampling_step[e] = next;
      if (C.limit_ang_vel_at_zero_command_prob > 0 && u01(r1.y) < C.limit_ang_vel_at_zero_command_prob) {
        cmd[2] = (u01(r1.z) < 0.5f) ? rng[4] : rng[5];
        if (heading) stop_heading[e] = 1;                          // :581-582
      }
    }
  }
  if (C.turn_over && GO2_EXT_PTR(const float*, &C, ext_turn_over_timer)[e] > 0) {   // turn-over zero-command time, legged_robot.py:585-590
    cmd[0] = 0; cmd[1] = 0; cmd[2] = 0;
    if (stop_heading) stop_heading[e] = 1;
  }
  acc[0] += cmd[0]; acc[1] += cmd[1];
}

// yaw-rate command from the heading target (legged_robot.py:411-419): quat_apply(base_quat, [1,0,0]), atan2, wrap_to_pi (math.py:15-18), clip
static void heading_to_yaw(const Go2EnvBuffers& B, int e) {
  const float* rs = B.root_states + (size_t)e * 13;
  float* cmd = B.commands + (size_t)e * GO2_NUM_CMD;
  const float* rng = B.env_command_ranges + (size_t)e * 6;
  const float qx = rs[3], qy = rs[4], qz = rs[5], qw = rs[6];
  const float ty = qz * 2.0f, tz = -qy * 2.0f;                           // t = 2 (q_xyz x [1,0,0])
  const float fx = 1.0f + (qy * tz - qz * ty);
  const float fy = qw * ty + (qz * 0.0f - qx * tz);
  const float hd = std::atan2(fy, fx);
  float a = std::fmod(cmd[3] - hd, 6.2831855f);
  if (a != 0.0f && a < 0.0f) a += 6.2831855f;
  if (a > 3.1415927f) a -= 6.2831855f;
  cmd[2] = std::min(std::max(0.5f * a, rng[4]), rng[5]);
}

static void reset_env(const Go2EnvConfig& C, const Go2Model& M, const Go2EnvBuffers& B, const Go2StepParams& sp, int e,
                      bool initial) {
  // legged_robot.py:180-245 (+ :620-707, :1143-1169)
  uint32_t ge = (uint32_t)(C.env_offset + e);
  for (int j = 0; j < GO2_NUM_DOF; ++j) {
    U4 r = philox4x32_10(ge, sp.common_step_counter, ST_RESET_DR, (uint32_t)j, C.seed_lo, C.seed_hi);
    size_t o = (size_t)e * GO2_NUM_DOF + j;
    if (C.randomize_motor_strength) B.motor_strengths[o] = C.motor_strength_range[1] * u01(r.x) + C.motor_strength_range[0];
    if (C.randomize_motor_zero_offset) B.motor_zero_offsets[o] = C.motor_zero_offset_range[1] * u01(r.y) + C.motor_zero_offset_range[0];
    if (C.randomize_pd_gains) {
      B.p_gains_multiplier[o] = C.kp_mult_range[1] * u01(r.z) + C.kp_mult_range[0];
      B.d_gains_multiplier[o] = C.kd_mult_range[1] * u01(r.w) + C.kd_mult_range[0];
    }
  }
  U4 s3 = philox4x32_10(ge, sp.common_step_counter, ST_RESET_STATE, 3, C.seed_lo, C.seed_hi);
  U4 s4 = philox4x32_10(ge, sp.common_step_counter, ST_RESET_STATE, 4, C.seed_lo, C.seed_hi);
  U4 s5 = philox4x32_10(ge, sp.common_step_counter, ST_RESET_STATE, 5, C.seed_lo, C.seed_hi);
  if (C.terrain_curriculum && !initial && C.mesh_type != 0) {  // _update_terrain_curriculum
    float dist = B.max_move_distance[e];
    const float* acc = B.commands_xy_accumulation + (size_t)e * 2;
    const float* cmd = B.commands + (size_t)e * GO2_NUM_CMD;
    bool up = dist > C.terrain_length / 2;
    bool down;
    if (C.move_down_by_accumulated_xy_command)
      down = (dist < std::sqrt(acc[0] * acc[0] + acc[1] * acc[1]) * (C.resampling_time * (1 - sp.zero_command_proba)) * 0.5f) && !up;
    else
      down = (dist < std::sqrt(cmd[0] * cmd[0] + cmd[1] * cmd[1]) * C.max_episode_length_s * 0.5f) && !up;
    int lvl = B.terrain_levels[e] + (up ? 1 : 0) - (down ? 1 : 0);
    if (lvl >= C.num_levels) lvl = (int)(s3.w % (uint32_t)C.num_levels);
    else lvl = std::max(lvl, 0);
    B.terrain_levels[e] = lvl;
    const float* org = B.terrain_origins + ((size_t)lvl * C.num_types + B.terrain_types[e]) * 3;
    for (int i = 0; i < 3; ++i) B.env_origins[(size_t)e * 3 + i] = org[i];
    B.max_move_distance[e] = 0;
  }
  for (int j = 0; j < GO2_NUM_DOF; ++j) {  // _reset_dofs
    U4 r = philox4x32_10(ge, sp.common_step_counter, ST_RESET_STATE, (uint32_t)(j / 4), C.seed_lo, C.seed_hi);
    uint32_t w = (j % 4 == 0) ? r.x : (j % 4 == 1) ? r.y : (j % 4 == 2) ? r.z : r.w;
    B.dof_pos[(size_t)e * GO2_NUM_DOF + j] = C.default_dof_pos[j] * (u01(w) + 0.5f);
    B.dof_vel[(size_t)e * GO2_NUM_DOF + j] = 0;
  }
  float* rs = B.root_states + (size_t)e * 13;  // _reset_root_states
  float yaw = (2 * 3.14159265358979323846f) * u01(s3.x) - 3.14159265358979323846f;
  for (int i = 0; i < 13; ++i) rs[i] = C.base_init_state[i];
  rs[3] = 0; rs[4] = 0; rs[5] = std::sin(yaw * 0.5f); rs[6] = std::cos(yaw * 0.5f);
  if (C.turn_over) {   // legged_robot.py:642-691: flipped initial poses; the masks compare a float32 draw with cumulative proportions
    float* tt = GO2_EXT_PTR(float*, &C, ext_turn_over_timer) + e;
    *tt = 0;
    float rp = u01(s5.z);
    bool back = rp >= 0 && rp < C.turn_over_proportions[0], side = rp >= C.turn_over_proportions[0] && rp < C.turn_over_proportions[1];
    bool none = rp >= C.turn_over_proportions[1] && rp < C.turn_over_proportions[2];
    if (back || side) {
      U4 s6 = philox4x32_10(ge, sp.common_step_counter, ST_RESET_STATE, 6, C.seed_lo, C.seed_hi);
      const float PI_F = 3.14159265358979323846f;
      float roll;
      if (back) { rs[2] = C.turn_over_back_height[1] * u01(s6.x) + C.turn_over_back_height[0]; roll = PI_F; *tt = C.turn_over_zero_time_back; }
      else { rs[2] = C.turn_over_side_height[1] * u01(s6.y) + C.turn_over_side_height[0]; roll = (u01(s5.w) < 0.5f) ? 0.5f * PI_F : -0.5f * PI_F; *tt = C.turn_over_zero_time_side; }
      float cy = std::cos(yaw * 0.5f), sy = std::sin(yaw * 0.5f), cr = std::cos(roll * 0.5f), sr = std::sin(roll * 0.5f);   // quat_from_euler_xyz(roll, 0, yaw)
      rs[6] = cy * cr; rs[3] = cy * sr; rs[4] = sy * sr; rs[5] = sy * cr;
    } else if (!none) {
      for (int i = 3; i < 7; ++i) rs[i] = C.base_init_state[i];
    }
  }
  for (int i = 0; i < 3; ++i) rs[i] += B.env_origins[(size_t)e * 3 + i];
  if (C.custom_origins) { rs[0] += 2 * u01(s3.y) - 1; rs[1] += 2 * u01(s3.z) - 1; }
  rs[7] = u01(s4.x) - 0.5f; rs[8] = u01(s4.y) - 0.5f; rs[9] = u01(s4.z) - 0.5f;
  rs[10] = u01(s4.w) - 0.5f; rs[11] = u01(s5.x) - 0.5f; rs[12] = u01(s5.y) - 0.5f;
  for (int j = 0; j < GO2_NUM_DOF; ++j) {
    size_t o = (size_t)e * GO2_NUM_DOF + j;
    B.actions[o] = 0; B.last_actions[o] = 0; B.last_dof_vel[o] = 0;
  }
  B.episode_length_buf[e] = 0;
  B.reset_buf[e] = 1;
  B.commands_resampling_step[e] = C.resampling_time / C.dt;
  B.commands_xy_accumulation[(size_t)e * 2] = 0;
  B.commands_xy_accumulation[(size_t)e * 2 + 1] = 0;
  resample_commands(C, B, sp, e, ST_CMD_RESET);
}

static void measure_heights(const Go2EnvConfig& C, const Go2EnvBuffers& B, int e) {
  // legged_robot.py:1188-1224 + math.py:8-12
  float* mh = B.measured_heights + (size_t)e * GO2_NUM_HEIGHT;
  if (C.mesh_type == 0) { for (int i = 0; i < GO2_NUM_HEIGHT; ++i) mh[i] = 0; return; }
  const float* rs = B.root_states + (size_t)e * 13;
  float qz = rs[5], qw = rs[6];
  float nrm = std::max(std::sqrt(qz * qz + qw * qw), 1e-9f);
  qz /= nrm; qw /= nrm;
  for (int i = 0; i < GO2_NUM_HEIGHT; ++i) {
    float bx = C.height_points[i][0], by = C.height_points[i][1];
    // quat_apply((0,0,qz,qw), (bx,by,0)): t = 2 * cross(xyz, b); out = b + w t + cross(xyz, t)
    float tx = 2 * (-qz * by), ty = 2 * (qz * bx);
    float px = bx + qw * tx + (-qz * ty), py = by + qw * ty + (qz * tx);
    px += rs[0]; py += rs[1];
    px += C.border; py += C.border;
    long ix = (long)(px / C.hscale), iy = (long)(py / C.hscale);   // .long() truncates toward zero
    ix = std::min(std::max(ix, 0L), (long)C.hf_rows - 2);
    iy = std::min(std::max(iy, 0L), (long)C.hf_cols - 2);
    int16_t h1 = B.height_samples[ix * C.hf_cols + iy], h2 = B.height_samples[(ix + 1) * C.hf_cols + iy],
            h3 = B.height_samples[ix * C.hf_cols + iy + 1];
    mh[i] = (float)std::min(std::min(h1, h2), h3) * C.vscale;
  }
}

static float dynamic_sigma(const Go2EnvConfig& C, const Go2EnvBuffers& B, int e, float tv, float vmin, float vmax) {
  // legged_robot.py:1300-1320
  float d = C.tracking_sigma;
  if (!C.terrain_curriculum || !C.dynamic_sigma || C.mesh_type == 0) return d;
  float target = C.ds_max_sigma[B.terrain_ids[e]];
  float sigma = d;
  if (tv >= vmin && tv < vmax) sigma = d + (tv - vmin) / (vmax - vmin) * (target - d);
  if (tv >= vmax) sigma = target;
  float ls = std::min(std::exp(((float)B.terrain_levels[e] + 1.0f) / 10.0f) - 1.0f, 1.0f);
  return d + ls * (sigma - d);
}

}  // namespace

// ================================================================================================ C API
extern "C" {

// one full LeggedRobot.step for every env (legged_robot.py:60-100); buffers are HOST pointers
int go2_oracle_step(const Go2EnvConfig* Cp, const Go2Model* Mp, const Go2EnvBuffers* Bp, const float* actions_in,
                    const Go2StepParams* spp) {
  const Go2EnvConfig& C = *Cp; const Go2Model& M = *Mp; const Go2EnvBuffers& B = *Bp; const Go2StepParams& sp = *spp;
  const int N = C.num_envs;
  Terrain T{Cp, B.height_samples};
  std::vector<double> acc(GO2_EP_STATS, 0.0), xacc(GO2_NUM_XREW, 0.0);
  std::vector<double> lvl_sum(9, 0.0), lvl_cnt(9, 0.0);
  int n_reset = 0;
  std::vector<char> bad(N, 0);
#pragma omp parallel for schedule(static)
  for (int e = 0; e < N; ++e) {
    uint32_t ge = (uint32_t)(C.env_offset + e);
    float* act = B.actions + (size_t)e * GO2_NUM_DOF;
    float* lact = B.last_actions + (size_t)e * GO2_NUM_DOF;
    float* tq = B.torques + (size_t)e * GO2_NUM_DOF;
    float* rs = B.root_states + (size_t)e * 13;
    for (int j = 0; j < GO2_NUM_DOF; ++j) act[j] = std::min(std::max(actions_in[(size_t)e * GO2_NUM_DOF + j], -C.clip_actions), C.clip_actions);
    int start = 0;
    if (C.randomize_action_delay) start = (int)(philox4x32_10(ge, sp.common_step_counter, ST_DELAY, 0, C.seed_lo, C.seed_hi).x % (uint32_t)(C.decimation + 1));
    real pos[3] = {rs[0], rs[1], rs[2]}, quat[4] = {rs[3], rs[4], rs[5], rs[6]}, lw[3] = {rs[7], rs[8], rs[9]}, aw[3] = {rs[10], rs[11], rs[12]};
    real q[GO2_NUM_DOF], qd[GO2_NUM_DOF], tau[GO2_NUM_DOF];
    for (int j = 0; j < GO2_NUM_DOF; ++j) { q[j] = B.dof_pos[(size_t)e * GO2_NUM_DOF + j]; qd[j] = B.dof_vel[(size_t)e * GO2_NUM_DOF + j]; }
    SubstepOut so;
    for (int s = 0; s < C.decimation; ++s) {
      for (int j = 0; j < GO2_NUM_DOF; ++j) {  // _compute_torques, legged_robot.py:594-618 ('P' control)
        size_t o = (size_t)e * GO2_NUM_DOF + j;
        float a_in = (C.randomize_action_delay && s < start) ? lact[j] : act[j];
        float kp = C.kp[j] * B.p_gains_multiplier[o], kd = C.kd[j] * B.d_gains_multiplier[o];
        float t = kp * (a_in * C.action_scale + C.default_dof_pos[j] - (float)q[j] + B.motor_zero_offsets[o]) - kd * (float)qd[j];
        if (C.control_type == 1) t = kp * (a_in * C.action_scale - (float)qd[j]) - kd * ((float)qd[j] - B.last_dof_vel[o]) / C.sim_dt;   // 'V', legged_robot.py:612-613
        else if (C.control_type == 2) t = a_in * C.action_scale;                                                                     // 'T', :614-615
        t = std::min(std::max(t, -M.effort[j]), M.effort[j]);
        if (C.randomize_motor_strength) t *= B.motor_strengths[o];
        tq[j] = t;                                                   // what the reference reports (legged_robot.py:79-81)
        tau[j] = std::min(std::max(t, -M.effort[j]), M.effort[j]);   // PhysX effort clamp (spec)
      }
      physics_substep(C, M, T, B.body_inertia + (size_t)e * GO2_NUM_DYN * GO2_INERTIA_STRIDE, B.friction_coeffs[e],
                      B.restitutions[e], pos, quat, lw, aw, q, qd, tau, so);
    }
    for (int i = 0; i < 3; ++i) { rs[i] = (float)pos[i]; rs[7 + i] = (float)lw[i]; rs[10 + i] = (float)aw[i]; }
    for (int i = 0; i < 4; ++i) rs[3 + i] = (float)quat[i];
    for (int j = 0; j < GO2_NUM_DOF; ++j) { B.dof_pos[(size_t)e * GO2_NUM_DOF + j] = (float)q[j]; B.dof_vel[(size_t)e * GO2_NUM_DOF + j] = (float)qd[j]; }
    for (int b = 0; b < GO2_NUM_REPORT; ++b) for (int i = 0; i < 3; ++i) B.contact_forces[((size_t)e * GO2_NUM_REPORT + b) * 3 + i] = (float)so.contact_force[b][i];
    bad[e] = 0;
    if (C.state_guard) {   // Go2EnvConfig.state_guard: non-finite state -> initial pose at the origin, zero velocities / torques / contacts, forced reset
      bool nf = false;
      for (int i = 0; i < 13; ++i) nf |= !(std::fabs(rs[i]) <= 3.0e38f);
      for (int j = 0; j < GO2_NUM_DOF; ++j) nf |= !(std::fabs(B.dof_pos[(size_t)e * GO2_NUM_DOF + j]) <= 3.0e38f) | !(std::fabs(B.dof_vel[(size_t)e * GO2_NUM_DOF + j]) <= 3.0e38f);
      if (nf) {
        bad[e] = 1;
        for (int i = 0; i < 3; ++i) { rs[i] = B.env_origins[(size_t)e * 3 + i] + C.base_init_state[i]; pos[i] = rs[i]; }
        for (int i = 3; i < 7; ++i) { rs[i] = C.base_init_state[i]; quat[i - 3] = rs[i]; }
        for (int i = 7; i < 13; ++i) rs[i] = 0;
        for (int i = 0; i < 3; ++i) lw[i] = aw[i] = 0;
        for (int j = 0; j < GO2_NUM_DOF; ++j) {
          size_t o = (size_t)e * GO2_NUM_DOF + j;
          B.dof_pos[o] = C.default_dof_pos[j]; B.dof_vel[o] = 0; B.last_dof_vel[o] = 0; tq[j] = 0; q[j] = C.default_dof_pos[j]; qd[j] = 0;
        }
        for (int i = 0; i < GO2_NUM_REPORT * 3; ++i) B.contact_forces[(size_t)e * GO2_NUM_REPORT * 3 + i] = 0;
      }
    }
    {  // feet position / velocity at the new configuration (rigid_body_states refresh, legged_robot.py:109)
      Kin K;
      kinematics(M, pos, quat, q, K);
      V3 wb = mulT(K.Rw[0], V3{aw[0], aw[1], aw[2]}), vb = mulT(K.Rw[0], V3{lw[0], lw[1], lw[2]});
      V6 v[GO2_NUM_DYN];
      v[0] = mk6(wb, vb);
      for (int b = 1; b < GO2_NUM_DYN; ++b) { v[b] = mul(K.X[b], v[parent_of(M, b)]); v[b].v[M.joint_axis[b - 1]] += qd[b - 1]; }
      for (int l = 0; l < 4; ++l) {
        int b = 3 + 3 * l;
        V3 r{(real)M.foot_offset[l][0], (real)M.foot_offset[l][1], (real)M.foot_offset[l][2]};
        V3 pw = K.pw[b] + mul(K.Rw[b], r);
        V3 vw = mul(K.Rw[b], lin(v[b]) + cross(ang(v[b]), r));
        float* fp = B.feet_pos + ((size_t)e * 4 + l) * 3; float* fv = B.feet_vel + ((size_t)e * 4 + l) * 3;
        fp[0] = (float)pw.x; fp[1] = (float)pw.y; fp[2] = (float)pw.z; fv[0] = (float)vw.x; fv[1] = (float)vw.y; fv[2] = (float)vw.z;
      }
    }
  }
  // ---------------- post_physics_step (legged_robot.py:102-142); serial (cross-env sums are order sensitive)
  for (int e = 0; e < N; ++e) {
    uint32_t ge = (uint32_t)(C.env_offset + e);
    float* rs = B.root_states + (size_t)e * 13;
    B.episode_length_buf[e] += 1;
    B.commands_resampling_step[e] -= 1;
    if (C.turn_over) { float* tt = GO2_EXT_PTR(float*, &C, ext_turn_over_timer) + e; *tt = std::max(*tt - C.dt, 0.0f); }   // legged_robot.py:114-115
    real qr[4] = {rs[3], rs[4], rs[5], rs[6]}, o3[3];
    real lv[3] = {rs[7], rs[8], rs[9]}, av[3] = {rs[10], rs[11], rs[12]}, gv[3] = {0, 0, -1};
    float* blv = B.base_lin_vel + (size_t)e * 3; float* bav = B.base_ang_vel + (size_t)e * 3; float* pg = B.projected_gravity + (size_t)e * 3;
    quat_rotate_inverse(qr, lv, o3); for (int i = 0; i < 3; ++i) blv[i] = (float)o3[i];
    quat_rotate_inverse(qr, av, o3); for (int i = 0; i < 3; ++i) bav[i] = (float)o3[i];
    quat_rotate_inverse(qr, gv, o3); for (int i = 0; i < 3; ++i) pg[i] = (float)o3[i];
    {
      float dx = rs[0] - B.env_origins[(size_t)e * 3], dy = rs[1] - B.env_origins[(size_t)e * 3 + 1];
      B.max_move_distance[e] = std::max(B.max_move_distance[e], std::sqrt(dx * dx + dy * dy));
    }
    if (B.commands_resampling_step[e] <= 0.0f && B.episode_length_buf[e] < C.max_episode_length - 1) resample_commands(C, B, sp, e, ST_CMD_CB);
    if (C.heading_command && !GO2_EXT_PTR(const uint8_t*, &C, ext_stop_heading)[e]) heading_to_yaw(B, e);
    measure_heights(C, B, e);
    // check_termination, legged_robot.py:170-178
    const float* cf = B.contact_forces + (size_t)e * GO2_NUM_REPORT * 3;
    bool term = !C.turn_over && std::sqrt(cf[0] * cf[0] + cf[1] * cf[1] + cf[2] * cf[2]) > 1.0f;   // legged_robot.py:174-175
    bool tout = B.episode_length_buf[e] > C.max_episode_length;
    B.time_out_buf[e] = tout; B.reset_buf[e] = term || tout || bad[e];
    // compute_reward, legged_robot.py:247-274, terms in enum order
    const float* cmd = B.commands + (size_t)e * GO2_NUM_CMD;
    const float* q = B.dof_pos + (size_t)e * GO2_NUM_DOF; const float* qd = B.dof_vel + (size_t)e * GO2_NUM_DOF;
    const float* tq = B.torques + (size_t)e * GO2_NUM_DOF; const float* act = B.actions + (size_t)e * GO2_NUM_DOF;
    float* lact = B.last_actions + (size_t)e * GO2_NUM_DOF; float* llact = B.last_last_actions + (size_t)e * GO2_NUM_DOF;
    float* lqd = B.last_dof_vel + (size_t)e * GO2_NUM_DOF;
    const float* mh = B.measured_heights + (size_t)e * GO2_NUM_HEIGHT;
    float term_v[GO2_NUM_REW];
    {
      float sx = dynamic_sigma(C, B, e, std::fabs(cmd[0]), C.ds_min_lin, C.ds_max_lin), sy = dynamic_sigma(C, B, e, std::fabs(cmd[1]), C.ds_min_lin, C.ds_max_lin);
      float ex = (cmd[0] - blv[0]) * (cmd[0] - blv[0]), ey = (cmd[1] - blv[1]) * (cmd[1] - blv[1]);
      term_v[GO2_REW_TRACKING_LIN_VEL] = std::exp(-(ex / sx + ey / sy));
      float sa = dynamic_sigma(C, B, e, std::fabs(cmd[2]), C.ds_min_ang, C.ds_max_ang);
      term_v[GO2_REW_TRACKING_ANG_VEL] = std::exp(-((cmd[2] - bav[2]) * (cmd[2] - bav[2])) / sa);
    }
    term_v[GO2_REW_LIN_VEL_Z] = blv[2] * blv[2];
    term_v[GO2_REW_ANG_VEL_XY] = bav[0] * bav[0] + bav[1] * bav[1];
    float s_acc = 0, s_pow = 0, s_tq = 0, s_rate = 0, s_smooth = 0, s_lim = 0, s_hip = 0;
    for (int j = 0; j < GO2_NUM_DOF; ++j) {
      float a1 = (lqd[j] - qd[j]) / C.dt; s_acc += a1 * a1;
      s_pow += std::fabs(tq[j] * qd[j]);
      s_tq += tq[j] * tq[j];
      s_rate += (lact[j] - act[j]) * (lact[j] - act[j]);
      float sm = act[j] - 2 * lact[j] + llact[j]; s_smooth += sm * sm;
      s_lim += -std::min(q[j] - C.soft_dof_limit_lo[j], 0.0f) + std::max(q[j] - C.soft_dof_limit_hi[j], 0.0f);
      if (j % 3 == 0) s_hip += std::fabs(q[j] - C.default_dof_pos[j]);
    }
    for (int j = 0; j < GO2_NUM_DOF; ++j) llact[j] = lact[j];  // legged_robot.py:1378
    term_v[GO2_REW_DOF_ACC] = s_acc; term_v[GO2_REW_DOF_POWER] = s_pow; term_v[GO2_REW_TORQUES] = s_tq;
    term_v[GO2_REW_ACTION_RATE] = s_rate; term_v[GO2_REW_ACTION_SMOOTHNESS] = s_smooth; term_v[GO2_REW_DOF_POS_LIMITS] = s_lim;
    term_v[GO2_REW_HIP_TO_DEFAULT] = s_hip;
    float base_height;
    {  // _get_base_height, legged_robot.py:1387-1397
      float sh = 0;
      for (int i = 0; i < GO2_NUM_HEIGHT; ++i) sh += mh[i] * C.base_height_mask[i];
      base_height = rs[2] - sh / C.num_base_height_points;
    }
    term_v[GO2_REW_CORRECT_BASE_HEIGHT] = (base_height - C.base_height_target) * (base_height - C.base_height_target);
    {  // collision: thigh + calf bodies (report ids 4,5 + 4l), legged_robot.py:1277-1279
      float cnt = 0;
      for (int l = 0; l < 4; ++l)
        for (int k = 1; k <= 2; ++k) {
          const float* f = cf + (3 + 4 * l + k) * 3;
          cnt += (std::sqrt(f[0] * f[0] + f[1] * f[1] + f[2] * f[2]) > 0.1f) ? 1.0f : 0.0f;
        }
      term_v[GO2_REW_COLLISION] = cnt;
    }
    {  // feet_regulation, legged_robot.py:1404-1414
      float r = 0;
      for (int l = 0; l < 4; ++l) {
        const float* fp = B.feet_pos + ((size_t)e * 4 + l) * 3; const float* fv = B.feet_vel + ((size_t)e * 4 + l) * 3;
        float f2b = (fp[0] - rs[0]) * pg[0] + (fp[1] - rs[1]) * pg[1] + (fp[2] - rs[2]) * pg[2];
        float fh = std::max(base_height - f2b, 0.0f);
        r += (fv[0] * fv[0] + fv[1] * fv[1]) * std::exp(-fh / (0.025f * C.base_height_target));
      }
      term_v[GO2_REW_FEET_REGULATION] = r;
    }
    float rew = 0;
    bool need_to = false;   // legged_robot.py:257-265; roll of get_euler_xyz (isaacgym_utils.py:11-17)
    if (C.turn_over) {
      float qx = rs[3], qy = rs[4], qz = rs[5], qw = rs[6];
      need_to = std::fabs(std::atan2(2.0f * (qw * qx + qy * qz), qw * qw - qx * qx - qy * qy + qz * qz)) > C.turn_over_roll_threshold;
    }
    for (int k = 0; k < GO2_NUM_REW; ++k) {
      float rk = term_v[k] * (need_to ? C.to_scales[k] : C.reward_scales[k]) * sp.reward_curriculum[k];
      rew += rk;
      B.episode_sums[(size_t)e * GO2_NUM_REW + k] += rk;
    }
    float term_rew = 0;
    if (C.num_xrew > 0) {
      // the reward functions no registered go2 task switches on (legged_robot.py:1236-1441, go2_env.py:62-68); a function with a zero scale is not
      // called by the reference (:920-938), so the stateful ones only advance when active
      float* xs = GO2_EXT_PTR(float*, &C, ext_xrew_sums) + (size_t)e * GO2_NUM_XREW;
      float* st = GO2_EXT_PTR(float*, &C, ext_xrew_state) + (size_t)e * 12;   // feet_air_time[4], last_contacts[4], last_contacts2[4]
      const float* sc = C.xrew_scales; const float* tsc = C.to_xscales;
      float xv[GO2_NUM_XREW] = {0};
      const float* feet_f[4]; const float* feet_p[4]; bool contact[4];
      for (int l = 0; l < 4; ++l) { feet_f[l] = cf + (6 + 4 * l) * 3; feet_p[l] = B.feet_pos + ((size_t)e * 4 + l) * 3; contact[l] = feet_f[l][2] > 1.0f; }
      const float cmd_xy = std::sqrt(cmd[0] * cmd[0] + cmd[1] * cmd[1]);
      xv[GO2_XREW_ORIENTATION] = pg[0] * pg[0] + pg[1] * pg[1];                                    // :1236-1238
      if (sc[GO2_XREW_BASE_HEIGHT] != 0 || tsc[GO2_XREW_BASE_HEIGHT] != 0) {                                                           // :1245-1259
        float nfc = 0, fcp[3] = {0, 0, 0};
        for (int l = 0; l < 4; ++l) {
          bool filt = contact[l] || st[8 + l] != 0;
          st[8 + l] = contact[l];
          if (filt) { nfc += 1; for (int k = 0; k < 3; ++k) fcp[k] += feet_p[l][k]; }
        }
        float den = std::max(nfc, 1.0f), bh = 0;
        for (int k = 0; k < 3; ++k) bh += (fcp[k] / den - rs[k]) * pg[k];
        xv[GO2_XREW_BASE_HEIGHT] = (bh - C.base_height_target) * (bh - C.base_height_target) * (nfc > 0 ? 1.0f : 0.0f);
      }
      float s_qd = 0, s_vl = 0, s_tl = 0, s_def = 0;
      for (int j = 0; j < GO2_NUM_DOF; ++j) {
        s_qd += qd[j] * qd[j];                                                                                             // dof_vel :1265-1267
        s_vl += std::min(std::max(std::fabs(qd[j]) - M.vel_limit[j] * C.soft_dof_vel_limit, 0.0f), 1.0f);                   // dof_vel_limits :1291-1294
        s_tl += std::max(std::fabs(tq[j]) - M.effort[j] * C.soft_torque_limit, 0.0f);                                      // torque_limits :1296-1298
        s_def += std::fabs(q[j] - C.default_dof_pos[j]);                                                                   // similar_to_default :1416-1418
      }
      xv[GO2_XREW_DOF_VEL] = s_qd; xv[GO2_XREW_DOF_VEL_LIMITS] = s_vl; xv[GO2_XREW_TORQUE_LIMITS] = s_tl; xv[GO2_XREW_SIMILAR_TO_DEFAULT] = s_def;
      xv[GO2_XREW_TERMINATION] = (B.reset_buf[e] && !B.time_out_buf[e]) ? 1.0f : 0.0f;               // :1281-1283
      if (sc[GO2_XREW_FEET_AIR_TIME] != 0 || tsc[GO2_XREW_FEET_AIR_TIME] != 0) {                                                          // :1347-1358
        float r = 0;
        for (int l = 0; l < 4; ++l) {
          bool filt = contact[l] || st[4 + l] != 0;
          st[4 + l] = contact[l];
          bool first = st[l] > 0 && filt;
          st[l] += C.dt;
          r += (st[l] - 0.5f) * (first ? 1.0f : 0.0f);
          if (filt) st[l] = 0;
        }
        xv[GO2_XREW_FEET_AIR_TIME] = r * (cmd_xy > 0.1f ? 1.0f : 0.0f);
      }
      bool stumble = false; float s_fc = 0;
      for (int l = 0; l < 4; ++l) {
        const float* f = feet_f[l];
        stumble = stumble || (std::sqrt(f[0] * f[0] + f[1] * f[1]) > 5 * std::fabs(f[2]));           // :1360-1363
        s_fc += std::max(std::sqrt(f[0] * f[0] + f[1] * f[1] + f[2] * f[2]) - C.max_contact_force, 0.0f);   // :1369-1371
      }
      xv[GO2_XREW_STUMBLE] = stumble; xv[GO2_XREW_FEET_CONTACT_FORCES] = s_fc;
      xv[GO2_XREW_STAND_STILL] = s_def * (cmd_xy < 0.1f ? 1.0f : 0.0f);                              // :1365-1367
      xv[GO2_XREW_UPRIGHT] = (-1 - pg[2]) / 2;                                                       // :1420-1421
      {                                                                                              // legs_distance :1423-1441
        float ly[4];
        for (int l = 0; l < 4; ++l) {
          real d[3] = {feet_p[l][0] - rs[0], feet_p[l][1] - rs[1], feet_p[l][2] - rs[2]}, o[3];
          quat_rotate_inverse(qr, d, o); ly[l] = (float)o[1];
        }
        float df = std::max(C.min_legs_distance - (ly[0] - ly[1]), 0.0f), dr = std::max(C.min_legs_distance - (ly[2] - ly[3]), 0.0f);
        xv[GO2_XREW_LEGS_DISTANCE] = df * df + dr * dr;
      }
      if (sc[GO2_XREW_X_COMMAND_HIP_REGULAR] != 0 || tsc[GO2_XREW_X_COMMAND_HIP_REGULAR] != 0)                                                  // go2_env.py:62-68
        xv[GO2_XREW_X_COMMAND_HIP_REGULAR] = (std::fabs(q[0] + q[3]) + std::fabs(q[6] + q[9])) * (std::fabs(cmd[0]) / std::sqrt(cmd[0] * cmd[0] + cmd[1] * cmd[1] + cmd[2] * cmd[2]));
      for (int k = 0; k < GO2_NUM_XREW; ++k) {
        if (sc[k] == 0 && tsc[k] == 0) continue;
        float rk = xv[k] * ((need_to && k != GO2_XREW_TERMINATION) ? tsc[k] : sc[k]) * sp.xrew_curriculum[k];
        if (k == GO2_XREW_TERMINATION) term_rew = rk; else rew += rk;
        xs[k] += rk;
      }
    }
    if (C.only_positive_rewards) rew = std::max(rew, 0.0f);   // legged_robot.py:266-267 (episode sums keep the unclipped terms)
    B.rew_buf[e] = rew + term_rew;                            // termination reward after the clip (:268-272)
    if (B.reset_buf[e]) {
      n_reset++;
      if (C.num_xrew > 0) {
        float* xs = GO2_EXT_PTR(float*, &C, ext_xrew_sums) + (size_t)e * GO2_NUM_XREW;
        for (int k = 0; k < GO2_NUM_XREW; ++k) { xacc[k] += xs[k]; xs[k] = 0; }
        for (int l = 0; l < 4; ++l) GO2_EXT_PTR(float*, &C, ext_xrew_state)[(size_t)e * 12 + l] = 0;    // feet_air_time, legged_robot.py:220
      }
      for (int k = 0; k < GO2_NUM_REW; ++k) { acc[k] += B.episode_sums[(size_t)e * GO2_NUM_REW + k]; B.episode_sums[(size_t)e * GO2_NUM_REW + k] = 0; }
      reset_env(C, M, B, sp, e, false);
    }
    if (C.push_robots && (B.episode_length_buf[e] % C.push_interval == 0)) {  // legged_robot.py:709-724
      U4 p0 = philox4x32_10(ge, sp.common_step_counter, ST_PUSH, 0, C.seed_lo, C.seed_hi);
      U4 p1 = philox4x32_10(ge, sp.common_step_counter, ST_PUSH, 1, C.seed_lo, C.seed_hi);
      // torch_rand_float(-m, m): (m - (-m)) * u + (-m); doubling is exact in binary floating point
      float sv = 2 * C.max_push_vel_xy, sa = 2 * C.max_push_ang_vel;
      rs[7] = sv * u01(p0.x) - C.max_push_vel_xy; rs[8] = sv * u01(p0.y) - C.max_push_vel_xy;
      rs[10] = sa * u01(p0.z) - C.max_push_ang_vel; rs[11] = sa * u01(p0.w) - C.max_push_ang_vel;
      rs[12] = sa * u01(p1.x) - C.max_push_ang_vel;
    }
    // compute_observations, go2_env.py:23-53 (q, qd, actions re-read: a reset may have changed them)
    float* ob = B.obs_buf + (size_t)e * GO2_NUM_OBS; float* pv = B.privileged_obs_buf + (size_t)e * GO2_NUM_PRIV;
    float tmp[GO2_NUM_OBS];
    for (int i = 0; i < 3; ++i) { tmp[i] = bav[i] * C.obs_scale_ang_vel; tmp[3 + i] = pg[i]; }
    tmp[6] = cmd[0] * C.obs_scale_lin_vel; tmp[7] = cmd[1] * C.obs_scale_lin_vel; tmp[8] = cmd[2] * C.obs_scale_ang_vel;
    for (int j = 0; j < GO2_NUM_DOF; ++j) {
      tmp[9 + j] = (q[j] - C.default_dof_pos[j]) * C.obs_scale_dof_pos;
      tmp[21 + j] = qd[j] * C.obs_scale_dof_vel;
      tmp[33 + j] = act[j];
    }
    for (int i = 0; i < 3; ++i) pv[i] = blv[i] * C.obs_scale_lin_vel;
    for (int i = 0; i < GO2_NUM_OBS; ++i) pv[3 + i] = tmp[i];
    for (int l = 0; l < 4; ++l) { const float* f = cf + (6 + 4 * l) * 3; pv[48 + l] = std::sqrt(f[0] * f[0] + f[1] * f[1] + f[2] * f[2]) * 1e-3f; }
    for (int j = 0; j < GO2_NUM_DOF; ++j) { pv[52 + j] = tq[j] / M.effort[j]; pv[64 + j] = (lqd[j] - qd[j]) / C.dt * 1e-4f; }
    for (int i = 0; i < GO2_NUM_HEIGHT; ++i) pv[76 + i] = std::min(std::max(rs[2] - 0.5f - mh[i], -1.0f), 1.0f) * C.obs_scale_height;
    for (int i = 0; i < GO2_NUM_OBS; ++i) {
      float x = tmp[i];
      if (C.add_noise) {
        U4 r = philox4x32_10(ge, sp.common_step_counter, ST_NOISE, (uint32_t)(i / 4), C.seed_lo, C.seed_hi);
        uint32_t w = (i % 4 == 0) ? r.x : (i % 4 == 1) ? r.y : (i % 4 == 2) ? r.z : r.w;
        x += (2 * u01(w) - 1) * C.noise_scale_vec[i];
      }
      ob[i] = std::min(std::max(x, -C.clip_obs), C.clip_obs);
    }
    for (int i = 0; i < GO2_NUM_PRIV; ++i) pv[i] = std::min(std::max(pv[i], -C.clip_obs), C.clip_obs);
    for (int j = 0; j < GO2_NUM_DOF; ++j) { lact[j] = act[j]; lqd[j] = qd[j]; }
  }
  if (C.num_xrew > 0) {   // the extra terms' episode means (rows behind the 14 int64 accumulators the kernel uses; the oracle sums in double)
    float* xst = reinterpret_cast<float*>(GO2_EXT_PTR(long long*, &C, ext_xrew_log) + GO2_NUM_XREW);
    for (int k = 0; k < GO2_NUM_XREW; ++k)
      xst[(size_t)sp.ep_slot * GO2_NUM_XREW + k] = n_reset > 0 ? (float)(xacc[k] / n_reset) / C.max_episode_length_s
                                                               : xst[(size_t)((sp.ep_slot + GO2_EP_SLOTS - 1) % GO2_EP_SLOTS) * GO2_NUM_XREW + k];
  }
  // extras["episode"], legged_robot.py:229-242 — only refreshed when at least one env reset
  if (n_reset > 0 && B.ep_stats) {
    float* st = B.ep_stats + (size_t)sp.ep_slot * GO2_EP_STATS;
    for (int k = 0; k < GO2_NUM_REW; ++k) st[k] = (float)(acc[k] / n_reset) / C.max_episode_length_s;
    double all = 0;
    for (int e = 0; e < N; ++e) { all += B.terrain_levels[e]; lvl_sum[B.terrain_ids[e]] += B.terrain_levels[e]; lvl_cnt[B.terrain_ids[e]] += 1; }
    st[GO2_NUM_REW] = C.mesh_type == 0 ? 0.0f : (float)(all / N);
    for (int t = 0; t < 9; ++t) st[GO2_NUM_REW + 1 + t] = lvl_cnt[t] > 0 ? (float)(lvl_sum[t] / lvl_cnt[t]) : 0.0f;
    st[GO2_NUM_REW + 10] = (float)n_reset;
    st[GO2_NUM_REW + 11] = 1.0f;
  } else if (B.ep_stats) {   // no reset: the previous step's row is served again (include/go2_b200.h: GO2_EP_SLOTS)
    const float* prev = B.ep_stats + (size_t)((sp.ep_slot + GO2_EP_SLOTS - 1) % GO2_EP_SLOTS) * GO2_EP_STATS;
    float* st = B.ep_stats + (size_t)sp.ep_slot * GO2_EP_STATS;
    for (int k = 0; k < GO2_EP_STATS; ++k) st[k] = prev[k];
  }
  return 0;
}

// reset_idx(all) at construction (base_task.py:82-86: reset_idx then step(zeros) is done by the caller)
int go2_oracle_reset_all(const Go2EnvConfig* C, const Go2Model* M, const Go2EnvBuffers* B, const Go2StepParams* sp) {
  for (int e = 0; e < C->num_envs; ++e) {
    for (int k = 0; k < GO2_NUM_REW; ++k) B->episode_sums[(size_t)e * GO2_NUM_REW + k] = 0;
    if (C->num_xrew > 0) {
      for (int k = 0; k < GO2_NUM_XREW; ++k) GO2_EXT_PTR(float*, C, ext_xrew_sums)[(size_t)e * GO2_NUM_XREW + k] = 0;
      for (int l = 0; l < 4; ++l) GO2_EXT_PTR(float*, C, ext_xrew_state)[(size_t)e * 12 + l] = 0;
    }
    reset_env(*C, *M, *B, *sp, e, true);
  }
  return 0;
}

// n physics substeps with given torques held constant (dynamics parity in isolation); root/dof state in place
int go2_oracle_substeps(const Go2EnvConfig* C, const Go2Model* M, const Go2EnvBuffers* B, const float* tau_in, int n) {
  Terrain T{C, B->height_samples};
#pragma omp parallel for schedule(static)
  for (int e = 0; e < C->num_envs; ++e) {
    float* rs = B->root_states + (size_t)e * 13;
    real pos[3] = {rs[0], rs[1], rs[2]}, quat[4] = {rs[3], rs[4], rs[5], rs[6]}, lw[3] = {rs[7], rs[8], rs[9]}, aw[3] = {rs[10], rs[11], rs[12]};
    real q[GO2_NUM_DOF], qd[GO2_NUM_DOF], tau[GO2_NUM_DOF];
    for (int j = 0; j < GO2_NUM_DOF; ++j) { q[j] = B->dof_pos[(size_t)e * GO2_NUM_DOF + j]; qd[j] = B->dof_vel[(size_t)e * GO2_NUM_DOF + j]; tau[j] = std::min(std::max(tau_in[(size_t)e * GO2_NUM_DOF + j], -M->effort[j]), M->effort[j]); }
    SubstepOut so;
    for (int s = 0; s < n; ++s)
      physics_substep(*C, *M, T, B->body_inertia + (size_t)e * GO2_NUM_DYN * GO2_INERTIA_STRIDE, B->friction_coeffs[e], B->restitutions[e], pos, quat, lw, aw, q, qd, tau, so);
    for (int i = 0; i < 3; ++i) { rs[i] = (float)pos[i]; rs[7 + i] = (float)lw[i]; rs[10 + i] = (float)aw[i]; }
    for (int i = 0; i < 4; ++i) rs[3 + i] = (float)quat[i];
    for (int j = 0; j < GO2_NUM_DOF; ++j) { B->dof_pos[(size_t)e * GO2_NUM_DOF + j] = (float)q[j]; B->dof_vel[(size_t)e * GO2_NUM_DOF + j] = (float)qd[j]; }
    for (int b = 0; b < GO2_NUM_REPORT; ++b) for (int i = 0; i < 3; ++i) B->contact_forces[((size_t)e * GO2_NUM_REPORT + b) * 3 + i] = (float)so.contact_force[b][i];
  }
  return 0;
}

// feet position / velocity of the current state (what refresh_rigid_body_state_tensor exposes for the 4 foot bodies)
int go2_oracle_feet(const Go2EnvConfig* C, const Go2Model* Mp, const Go2EnvBuffers* B) {
  const Go2Model& M = *Mp;
  for (int e = 0; e < C->num_envs; ++e) {
    const float* rs = B->root_states + (size_t)e * 13;
    real pos[3] = {rs[0], rs[1], rs[2]}, quat[4] = {rs[3], rs[4], rs[5], rs[6]}, q[GO2_NUM_DOF], qd[GO2_NUM_DOF];
    for (int j = 0; j < GO2_NUM_DOF; ++j) { q[j] = B->dof_pos[(size_t)e * GO2_NUM_DOF + j]; qd[j] = B->dof_vel[(size_t)e * GO2_NUM_DOF + j]; }
    Kin K;
    kinematics(M, pos, quat, q, K);
    V3 wb = mulT(K.Rw[0], V3{rs[10], rs[11], rs[12]}), vb = mulT(K.Rw[0], V3{rs[7], rs[8], rs[9]});
    V6 v[GO2_NUM_DYN];
    v[0] = mk6(wb, vb);
    for (int b = 1; b < GO2_NUM_DYN; ++b) { v[b] = mul(K.X[b], v[parent_of(M, b)]); v[b].v[M.joint_axis[b - 1]] += qd[b - 1]; }
    for (int l = 0; l < 4; ++l) {
      int b = 3 + 3 * l;
      V3 r{(real)M.foot_offset[l][0], (real)M.foot_offset[l][1], (real)M.foot_offset[l][2]};
      V3 pw = K.pw[b] + mul(K.Rw[b], r);
      V3 vw = mul(K.Rw[b], lin(v[b]) + cross(ang(v[b]), r));
      float* fp = B->feet_pos + ((size_t)e * 4 + l) * 3; float* fv = B->feet_vel + ((size_t)e * 4 + l) * 3;
      fp[0] = (float)pw.x; fp[1] = (float)pw.y; fp[2] = (float)pw.z; fv[0] = (float)vw.x; fv[1] = (float)vw.y; fv[2] = (float)vw.z;
    }
  }
  return 0;
}

// Philox known-answer hook (tests pin it against the Random123 reference vectors)
void go2_oracle_philox(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t* out) {
  U4 r = philox4x32_10(c0, c1, c2, c3, k0, k1);
  out[0] = r.x; out[1] = r.y; out[2] = r.z; out[3] = r.w;
}

int go2_oracle_sizeof_real(void) { return (int)sizeof(real); }
}

"""Host stand-in for the TRAINER half of libgo2b200.so (include/go2_b200.h, "PPO trainer kernels" and "Concurrent Teacher-Student" sections).

TEST INFRASTRUCTURE ONLY.  Every entry point of the RL section is restated in numpy (fp32 arithmetic, the kernels' formulas from
csrc/rl_kernels.cu / csrc/cts_kernels.cu / the host launchers of csrc/gemm_tc.cu, INCLUDING their argument checks: TMA operands must be
16-byte aligned with a 16-byte multiple pitch, narrow-head kernels take N <= 16 / K <= 128, workspaces must be large enough).  `install()`
points go2_rl_gym_b200.rl._ops at it so that the HOST logic above the C ABI (PPO / CTS / MoE-CTS algorithm classes, MlpEngine, flat
parameter storage, mini-batch bookkeeping) runs on CPU tensors and can be compared with the reference-made fixtures in the
`-m "not gpu"` suite.  The product never imports this file: without CUDA + the library `_ops.lib()` raises.
"""
import ctypes as C

import numpy as np

F = np.float32
_HALF_LOG_2PI = F(0.9189385332046727)


class EmuError(RuntimeError):
    pass


def _need(cond, msg):
    if not cond:
        raise EmuError(msg)


def _vec(ptr, n, ctype=C.c_float):
    if n <= 0:
        return np.zeros(0, dtype=np.dtype(ctype))
    _need(ptr, "null pointer")
    return np.ctypeslib.as_array((ctype * int(n)).from_address(int(ptr)))


def _mat(ptr, rows, cols, ld, ctype=C.c_float):
    """[rows, cols] view with leading dimension ld (elements)."""
    rows, cols, ld = int(rows), int(cols), int(ld)
    if rows == 0 or cols == 0:
        return np.zeros((rows, cols), dtype=F)
    _need(ld >= cols, f"leading dimension {ld} < {cols} columns")
    flat = _vec(ptr, (rows - 1) * ld + cols, ctype)
    return np.lib.stride_tricks.as_strided(flat, shape=(rows, cols), strides=(ld * flat.itemsize, flat.itemsize))


def _tma(ptr, ld, what):
    _need(int(ptr) % 16 == 0 and int(ld) % 4 == 0, f"{what}: TMA operand must be 16-byte aligned with a 16-byte multiple row pitch")


def _trunc_tf32(a):
    return (np.ascontiguousarray(a, dtype=F).view(np.uint32) & np.uint32(0xFFFFE000)).view(F)


def _rna_tf32(a):
    """cvt.rna.tf32.f32: round to nearest (ties away from zero) at 10 mantissa bits"""
    u = np.ascontiguousarray(a, dtype=F).view(np.uint32)
    return ((u + np.uint32(0x1000)) & np.uint32(0xFFFFE000)).view(F)


def _mm_tc(a, b):
    """a @ b.T as the tensor-core entry points form it (csrc/gemm_tc.cu), selected by GO2_EMU_TF32 (checks on the CPU that the bars of the GPU tests
    are realistic for new wirings): '0' (default) plain fp32; '1' one tf32 pass: operands lose their low 13 mantissa bits, like tcgen05 kind::tf32
    reading fp32 data; '3' the library's default 3xTF32 split: hi = trunc_tf32(a) (what the tensor core reads from the raw word), lo = rna_tf32(a - hi),
    lo_a hi_b + hi_a lo_b + hi_a hi_b; '3rw' the A/B variant go2_gemm_set_split(1): hi = rna_tf32(a), lo = a - hi (read truncated)."""
    import os
    mode = os.environ.get("GO2_EMU_TF32", "0")
    if mode == "1":
        return _trunc_tf32(a) @ _trunc_tf32(b).T
    if mode in ("3", "3rw"):
        if mode == "3":
            ah, bh = _trunc_tf32(a), _trunc_tf32(b)
            al, bl = _rna_tf32(np.asarray(a, F) - ah), _rna_tf32(np.asarray(b, F) - bh)
        else:
            ah, bh = _rna_tf32(a), _rna_tf32(b)
            al, bl = _trunc_tf32(np.asarray(a, F) - ah), _trunc_tf32(np.asarray(b, F) - bh)
        return ((al @ bh.T).astype(F) + (ah @ bl.T).astype(F)).astype(F) + (ah @ bh.T).astype(F)
    return a @ b.T


def _elu(x):
    return np.where(x > 0, x, np.expm1(np.minimum(x, 0)).astype(F)).astype(F)


def _elu_grad_from_out(y):      # ELU'(x) expressed with y = ELU(x): 1 for y > 0, y + 1 otherwise
    return np.where(y > 0, F(1), y + F(1)).astype(F)


# ---------------------------------------------------------------------------------------------------- Philox4x32-10
def philox(c0, c1, c2, c3, k0, k1):
    c = [np.asarray(x, dtype=np.uint64) & 0xFFFFFFFF for x in (c0, c1, c2, c3)]
    c = list(np.broadcast_arrays(*c))
    k0, k1 = int(k0) & 0xFFFFFFFF, int(k1) & 0xFFFFFFFF
    for _ in range(10):
        p0, p1 = np.uint64(0xD2511F53) * c[0], np.uint64(0xCD9E8D57) * c[2]
        h0, l0, h1, l1 = p0 >> np.uint64(32), p0 & 0xFFFFFFFF, p1 >> np.uint64(32), p1 & 0xFFFFFFFF
        c = [h1 ^ c[1] ^ np.uint64(k0), l1, h0 ^ c[3] ^ np.uint64(k1), l0]
        k0, k1 = (k0 + 0x9E3779B9) & 0xFFFFFFFF, (k1 + 0xBB67AE85) & 0xFFFFFFFF
    return [x.astype(np.uint32) for x in c]


def _u01(x):
    return (x >> np.uint32(8)).astype(F) * F(1.0 / 16777216.0)


# ---------------------------------------------------------------------------------------------------- GEMM family
def _linear_forward(X, ldx, W, ldw, b, Y, ldy, Yt, ldyt, M, N, K, act, tc):
    if tc:
        _tma(X, ldx, "forward_tc X"); _tma(W, ldw, "forward_tc W")
    x, w = _mat(X, M, K, ldx), _mat(W, N, K, ldw)
    y = _mm_tc(x, w) if tc else x @ w.T
    if b:
        y = y + _vec(b, N)
    y = y.astype(F)
    if act:
        y = _elu(y)
    if Y:
        _mat(Y, M, N, ldy)[...] = y
    if Yt:
        _mat(Yt, N, M, ldyt)[...] = y.T
    return 0


def go2_linear_forward_simt(X, ldx, W, ldw, b, Y, ldy, Yt, ldyt, M, N, K, act, stream):
    return _linear_forward(X, ldx, W, ldw, b, Y, ldy, Yt, ldyt, M, N, K, act, False)


def go2_linear_forward_tc(X, ldx, W, ldw, b, Y, ldy, Yt, ldyt, M, N, K, act, stream):
    return _linear_forward(X, ldx, W, ldw, b, Y, ldy, Yt, ldyt, M, N, K, act, True)


def go2_linear_dgrad_simt(dY, lddy, W, ldw, act_in, ldact, dX, lddx, dXt, lddxt, M, N, K, stream):
    d = (_mat(dY, M, N, lddy) @ _mat(W, N, K, ldw)).astype(F)
    if act_in:
        d = d * _elu_grad_from_out(_mat(act_in, M, K, ldact))
    _mat(dX, M, K, lddx)[...] = d
    if dXt:
        _mat(dXt, K, M, lddxt)[...] = d.T
    return 0


def go2_linear_dgrad_tc(dZ, lddz, Wt, ldwt, act_in, ldact, act_in_t, ldact_t, dX, lddx, dXt, lddxt, M, N, K, stream):
    _tma(dZ, lddz, "dgrad_tc dZ"); _tma(Wt, ldwt, "dgrad_tc Wt")
    d = _mm_tc(_mat(dZ, M, N, lddz), _mat(Wt, K, N, ldwt)).astype(F)
    if act_in:
        d = d * _elu_grad_from_out(_mat(act_in, M, K, ldact))
    elif act_in_t:
        d = d * _elu_grad_from_out(_mat(act_in_t, K, M, ldact_t).T)
    if dX:
        _mat(dX, M, K, lddx)[...] = d
    if dXt:
        _mat(dXt, K, M, lddxt)[...] = d.T
    return 0


def go2_linear_wgrad_simt(dY, lddy, X, ldx, dW, lddw, db, M, N, K, ws, wsn, stream):
    _need(not db, "go2_linear_wgrad_simt: bias gradient moved to go2_colsum")
    _mat(dW, N, K, lddw)[...] = (_mat(dY, M, N, lddy).T @ _mat(X, M, K, ldx)).astype(F)
    return 0


def go2_linear_wgrad_tc_rm(dZ, lddz, X, ldx, dW, lddw, db, M, N, K, ws, wsn, stream):
    _need(ws, "go2_linear_wgrad_tc_rm: workspace required")
    _tma(dZ, lddz, "wgrad_tc_rm dZ"); _tma(X, ldx, "wgrad_tc_rm X")
    kk = K + 1 if db else K
    ldp, rows_pad = (kk + 3) // 4 * 4, (N + 127) // 128 * 128
    _need(rows_pad * ldp <= wsn, "go2_linear_wgrad_tc_rm: workspace too small")
    _need(ldx >= kk, "go2_linear_wgrad_tc_rm: no room for the ones column")
    full = _mm_tc(_mat(dZ, M, N, lddz).T, _mat(X, M, kk, ldx).T).astype(F)
    _mat(dW, N, K, lddw)[...] = full[:, :K]
    if db:
        _vec(db, N)[...] = full[:, K]
    return 0


def go2_linear_wgrad_tc(dZt, lddzt, Xt, ldxt, dW, lddw, db, M, N, K, ws, wsn, stream):
    _tma(dZt, lddzt, "wgrad_tc dZt"); _tma(Xt, ldxt, "wgrad_tc Xt")
    kk = K + 1 if db else K
    full = _mm_tc(_mat(dZt, N, M, lddzt), _mat(Xt, kk, M, ldxt)).astype(F)
    _mat(dW, N, K, lddw)[...] = full[:, :K]
    if db:
        _vec(db, N)[...] = full[:, K]
    return 0


def _smalln_ok(name, N, K):
    _need(N <= 16 and K <= 128, f"{name}: needs N <= 16 and K <= 128")


def go2_linear_forward_smalln(X, ldx, W, ldw, b, Y, ldy, M, N, K, stream):
    _smalln_ok("go2_linear_forward_smalln", N, K)
    y = _mat(X, M, K, ldx) @ _mat(W, N, K, ldw).T
    if b:
        y = y + _vec(b, N)
    _mat(Y, M, N, ldy)[...] = y.astype(F)
    return 0


def go2_linear_dgrad_smalln(dY, lddy, W, ldw, act_in, ldact, dX, lddx, M, N, K, stream):
    _smalln_ok("go2_linear_dgrad_smalln", N, K)
    return go2_linear_dgrad_simt(dY, lddy, W, ldw, act_in, ldact, dX, lddx, 0, 0, M, N, K, stream)


def go2_linear_wgrad_smalln(dY, lddy, X, ldx, dW, lddw, db, M, N, K, ws, wsn, stream):
    _smalln_ok("go2_linear_wgrad_smalln", N, K)
    _need(ws and wsn >= N * K + N, "go2_linear_wgrad_smalln: workspace too small")
    dy = _mat(dY, M, N, lddy)
    _mat(dW, N, K, lddw)[...] = (dy.T @ _mat(X, M, K, ldx)).astype(F)
    if db:
        _vec(db, N)[...] = dy.sum(0, dtype=F)
    return 0


def go2_linear_wgrad_rank1(dY, lddy, X, ldx, dW, db, M, K, ws, wsn, stream):
    _need(K <= 512, "go2_linear_wgrad_rank1: K > 512")
    _need(ws and wsn >= K + 1, "go2_linear_wgrad_rank1: workspace too small")
    dy = _mat(dY, M, 1, lddy)
    _vec(dW, K)[...] = (dy.T @ _mat(X, M, K, ldx)).astype(F)[0]
    if db:
        _vec(db, 1)[...] = dy.sum(dtype=F)
    return 0


def go2_refresh_weights(n, src, ldin, dst, ldout, rows, cols, transpose, stream):
    _need(1 <= n <= 8, "go2_refresh_weights: 1..8 jobs, no null arrays")
    for j in range(n):
        _need(src[j] and dst[j] and rows[j] > 0 and cols[j] > 0, "go2_refresh_weights: bad job")
        s = _mat(src[j], rows[j], cols[j], ldin[j])
        if transpose[j]:
            _mat(dst[j], cols[j], rows[j], ldout[j])[...] = s.T
        else:
            _mat(dst[j], rows[j], cols[j], ldout[j])[...] = s
    return 0


def go2_transpose(inp, ldin, out, ldout, rows, cols, stream):
    _mat(out, cols, rows, ldout)[...] = _mat(inp, rows, cols, ldin).T
    return 0


def go2_colsum(dY, lddy, db, M, N, scratch, stream):
    _need(scratch, "go2_colsum: scratch required")
    _vec(db, N)[...] = _mat(dY, M, N, lddy).sum(0, dtype=F)
    return 0


# ---------------------------------------------------------------------------------------------------- rollout kernels
def _sample(mu, std_param, actions, logp, mu_out, sigma_out, N, A, seed, step, env_offset):
    m, s = _mat(mu, N, A, A), _vec(std_param, A)
    e = np.arange(N, dtype=np.uint64) + np.uint64(env_offset)
    z = np.zeros((N, (A + 3) // 4 * 4), dtype=F)
    two_pi = F(6.283185307179586)
    for b in range((A + 3) // 4):
        r = philox(e, step, 16, b, seed & 0xFFFFFFFF, seed >> 32)
        u0 = ((r[0] >> np.uint32(8)).astype(F) + F(1)) * F(1.0 / 16777216.0)
        u2 = ((r[2] >> np.uint32(8)).astype(F) + F(1)) * F(1.0 / 16777216.0)
        u1, u3 = _u01(r[1]), _u01(r[3])
        r0, r1 = np.sqrt(F(-2) * np.log(u0)).astype(F), np.sqrt(F(-2) * np.log(u2)).astype(F)
        z[:, 4 * b + 0], z[:, 4 * b + 1] = r0 * np.cos(two_pi * u1), r0 * np.sin(two_pi * u1)
        z[:, 4 * b + 2], z[:, 4 * b + 3] = r1 * np.cos(two_pi * u3), r1 * np.sin(two_pi * u3)
    a = (m + s * z[:, :A]).astype(F)
    _mat(actions, N, A, A)[...] = a
    _mat(mu_out, N, A, A)[...] = m
    _mat(sigma_out, N, A, A)[...] = np.broadcast_to(s, (N, A))
    d = a - m
    _vec(logp, N)[...] = (-(d * d) / (F(2) * s * s) - np.log(s) - _HALF_LOG_2PI).sum(1, dtype=F)
    return 0


def go2_sample_actions(mu, std_param, actions, logp, mu_out, sigma_out, N, A, seed, step, env_offset, stream):
    return _sample(mu, std_param, actions, logp, mu_out, sigma_out, N, A, int(seed), int(step) & 0xFFFFFFFF, env_offset)


def go2_sample_actions_dev(mu, std_param, actions, logp, mu_out, sigma_out, N, A, seed, d_step, env_offset, stream):
    _need(d_step, "go2_sample_actions_dev: null step pointer")
    return _sample(mu, std_param, actions, logp, mu_out, sigma_out, N, A, int(seed), int(_vec(d_step, 1, C.c_uint32)[0]), env_offset)


def go2_process_env_step(rew, dones, time_outs, values, rew_out, dones_out, N, gamma, perm, stream):
    src = _vec(perm, N, C.c_int64) if perm else np.arange(N)
    r = _vec(rew, N)[src].copy()
    if time_outs:
        r = r + F(gamma) * (_vec(values, N) * _vec(time_outs, N, C.c_uint8)[src].astype(F))
    _vec(rew_out, N)[...] = r.astype(F)
    _vec(dones_out, N, C.c_uint8)[...] = _vec(dones, N, C.c_uint8)[src]
    return 0


def go2_gae(rewards, values, dones, last_values, returns, advantages, T, N, gamma, lam, stats, stream):
    rw, v, dn = _mat(rewards, T, N, N), _mat(values, T, N, N), _mat(dones, T, N, N, C.c_uint8)
    ret, adv = _mat(returns, T, N, N), _mat(advantages, T, N, N)
    a, next_v = np.zeros(N, dtype=F), _vec(last_values, N).copy()
    g, l = F(gamma), F(lam)
    for t in range(T - 1, -1, -1):
        nt = F(1) - dn[t].astype(F)
        delta = rw[t] + nt * g * next_v - v[t]
        a = (delta + nt * g * l * a).astype(F)
        ret[t] = a + v[t]
        adv[t] = ret[t] - v[t]
        next_v = v[t].copy()
    st = _vec(stats, 2, C.c_double)
    st[0], st[1] = adv.astype(np.float64).sum(), (adv.astype(np.float64) ** 2).sum()
    return 0


def go2_adv_normalize(advantages, n, stats, global_count, stream):
    st, adv = _vec(stats, 2, C.c_double), _vec(advantages, n)
    mean = st[0] / global_count
    var = (st[1] - global_count * mean * mean) / (global_count - 1.0)
    sd = F(np.sqrt(max(var, 0.0)))
    adv[...] = (adv - F(mean)) / (sd + F(1e-8))
    return 0


def _pad_fill(out, w):
    """Padding columns of a gathered / concatenated row: zero, except column w (the "ones" column of the bias gradient)."""
    if out.shape[1] > w:
        out[:, w:] = 0
        out[:, w] = 1


def go2_gather_rows(src, width, idx, dst, ldd, dst_t, n, stream):
    rows = _vec(idx, n, C.c_int64) if idx else np.arange(n)
    hi = int(rows.max()) + 1 if n else 0
    s = _mat(src, hi, width, width)[rows]
    if dst:
        out = _mat(dst, n, ldd, ldd)
        out[:, :width] = s
        _pad_fill(out, width)
    if dst_t:
        _mat(dst_t, width, n, n)[...] = s.T
    return 0


def go2_gather_u8(src, perm, out, n, stream):
    p = _vec(perm, n, C.c_int64)
    _vec(out, n, C.c_uint8)[...] = _vec(src, int(p.max()) + 1, C.c_uint8)[p]
    return 0


# ---------------------------------------------------------------------------------------------------- PPO loss / optimiser
def go2_ppo_loss(mu, std_param, value, actions, old_logp, adv, target_values, returns, old_mu, old_sigma, dmu, dmu_t, dvalue, scal, M, A, clip,
                 value_coef, entropy_coef, use_clipped, inv_count, split, inv_a, inv_b, stream):
    _need(A <= 16, "go2_ppo_loss: at most 16 actions")
    clip, value_coef, entropy_coef, inv_count, inv_a, inv_b = map(F, (clip, value_coef, entropy_coef, inv_count, inv_a, inv_b))
    m, s, a = _mat(mu, M, A, A), _vec(std_param, A), _mat(actions, M, A, A)
    om, osig = _mat(old_mu, M, A, A), _mat(old_sigma, M, A, A)
    d = a - m
    lp = (-(d * d) / (F(2) * s * s) - np.log(s) - _HALF_LOG_2PI).sum(1, dtype=F)
    kl = (np.log(s / osig + F(1e-5)) + (osig * osig + (om - m) ** 2) / (F(2) * s * s) - F(0.5)).sum(1, dtype=F)
    ent = np.full(M, (F(0.5) + _HALF_LOG_2PI + np.log(s)).sum(dtype=F), dtype=F)
    A_ = _vec(adv, M)
    ratio = np.exp(lp - _vec(old_logp, M)).astype(F)
    s1, s2 = -A_ * ratio, -A_ * np.clip(ratio, F(1) - clip, F(1) + clip)
    surr = np.maximum(s1, s2)
    inside = (ratio > F(1) - clip) & (ratio < F(1) + clip)
    dlp = np.where(s1 >= s2, -A_ * ratio, np.where(inside, -A_ * ratio, F(0))).astype(F)
    rows = np.arange(M)
    dlp = dlp * np.where(rows < split, inv_a, inv_b)
    v, tv, ret = _vec(value, M), _vec(target_values, M), _vec(returns, M)
    if use_clipped:
        diff = v - tv
        vc = tv + np.clip(diff, -clip, clip)
        l1, l2 = (v - ret) ** 2, (vc - ret) ** 2
        vl = np.maximum(l1, l2)
        dv = np.where(l1 >= l2, F(2) * (v - ret), np.where((diff > -clip) & (diff < clip), F(2) * (vc - ret), F(0)))
    else:
        vl, dv = (ret - v) ** 2, F(2) * (v - ret)
    _vec(dvalue, M)[...] = (value_coef * dv * inv_count).astype(F)
    g = (dlp[:, None] * d / (s * s)).astype(F)
    _mat(dmu, M, A, A)[...] = g
    if dmu_t:
        _mat(dmu_t, A, M, M)[...] = g.T
    dstd = (dlp[:, None] * (d * d / (s * s * s) - F(1) / s) - entropy_coef * inv_count / s).sum(0, dtype=F)
    sc = _vec(scal, 20)
    sc[...] = 0
    sc[0], sc[2], sc[3] = kl.sum(dtype=F), vl.sum(dtype=F), ent.sum(dtype=F)
    sc[1], sc[19] = surr[rows < split].sum(dtype=F), surr[rows >= split].sum(dtype=F)
    sc[4:4 + A] = dstd
    return 0


def go2_kl_adaptive_lr(scal, count, desired_kl, lr_state, log_out, count_a, count_b, stream):
    sc, lrs = _vec(scal, 20), _vec(lr_state, 4)
    count, count_a, count_b, desired_kl = F(count), F(count_a), F(count_b), F(desired_kl)
    kl_mean = sc[0] / count
    lr = lrs[0]
    if desired_kl > 0:
        if kl_mean > desired_kl * F(2):
            lr = max(F(1e-5), lr / F(1.5))
        elif kl_mean < desired_kl / F(2) and kl_mean > 0:
            lr = min(F(1e-2), lr * F(1.5))
    lrs[0] = lr
    if log_out:
        lg = _vec(log_out, 5)
        lg[0] += sc[2] / count
        lg[1] += sc[1] / count_a + sc[19] / count_b
        lg[2], lg[3] = kl_mean, lr
        lg[4] += sc[3] / count
    return 0


def go2_adam_clip_step(params, grads, exp_avg, exp_avg_sq, n, max_norm, lr_state, grad_scale, scratch, stream):
    _need(scratch, "go2_adam_clip_step: scratch required")
    p, g, m, v, lrs = _vec(params, n), _vec(grads, n), _vec(exp_avg, n), _vec(exp_avg_sq, n), _vec(lr_state, 4)
    b1, b2, eps, grad_scale = F(0.9), F(0.999), F(1e-8), F(grad_scale)
    t = lrs[1] + F(1)
    lrs[1] = t
    lrs[2] = F(1.0 - 0.9 ** float(t))
    lrs[3] = F(np.sqrt(1.0 - 0.999 ** float(t)))
    total = np.sqrt((g * g).sum(dtype=F)).astype(F) * grad_scale
    coef = min(F(max_norm) / (total + F(1e-6)), F(1))
    gi = g * grad_scale * coef
    m[...] = b1 * m + (F(1) - b1) * gi
    v[...] = b2 * v + (F(1) - b2) * gi * gi
    p[...] = p - (lrs[0] / lrs[2]) * (m / (np.sqrt(v) / lrs[3] + eps))
    return 0


# ---------------------------------------------------------------------------------------------------- CTS / MoE pieces
def go2_concat2(a, wa, lda, b, wb, ldb, out, ld, out_t, n, stream):
    o = _mat(out, n, ld, ld)
    o[:, :wa] = _mat(a, n, wa, lda)
    o[:, wa:wa + wb] = _mat(b, n, wb, ldb)
    _pad_fill(o, wa + wb)
    if out_t:
        _mat(out_t, wa + wb, n, n)[...] = o[:, :wa + wb].T
    return 0


def go2_l2norm_forward(x, ldx, y, ldy, norm, n, d, stream):
    xv = _mat(x, n, d, ldx)
    nr = np.maximum(np.sqrt((xv * xv).sum(1, dtype=F)), F(1e-12)).astype(F)
    if norm:
        _vec(norm, n)[...] = nr
    _mat(y, n, d, ldy)[...] = xv * (F(1) / nr)[:, None]
    return 0


def go2_l2norm_backward(dy, lddy, y, ldy, norm, dx, lddx, dx_t, n, d, stream):
    dyv, yv = _mat(dy, n, d, lddy), _mat(y, n, d, ldy)
    dot = (yv * dyv).sum(1, dtype=F)
    g = ((dyv - yv * dot[:, None]) * (F(1) / _vec(norm, n))[:, None]).astype(F)
    _mat(dx, n, d, lddx)[...] = g
    if dx_t:
        _mat(dx_t, d, n, n)[...] = g.T
    return 0


def go2_grouped_linear_forward(X, ldx, W, b, Y, ldy, M, E, D, H, stream):
    _need(X and W and b and Y and M > 0 and E > 0 and D > 0 and H > 0, "go2_grouped_linear_forward: bad argument")
    x, y = _mat(X, M, E * H, ldx), _mat(Y, M, E * D, ldy)
    w, bias = _mat(W, E * D, H, H), _vec(b, E * D)
    for e in range(E):
        y[:, e * D:(e + 1) * D] = (x[:, e * H:(e + 1) * H] @ w[e * D:(e + 1) * D].T + bias[e * D:(e + 1) * D]).astype(F)
    return 0


def go2_grouped_linear_dgrad(dY, lddy, W, act, ldact, dX, lddx, M, E, D, H, stream):
    _need(dY and W and dX and M > 0 and E > 0 and D > 0 and H > 0, "go2_grouped_linear_dgrad: bad argument")
    g, dx, w = _mat(dY, M, E * D, lddy), _mat(dX, M, E * H, lddx), _mat(W, E * D, H, H)
    for e in range(E):
        d = (g[:, e * D:(e + 1) * D] @ w[e * D:(e + 1) * D]).astype(F)
        if act:
            d = d * _elu_grad_from_out(_mat(act, M, E * H, ldact)[:, e * H:(e + 1) * H])
        dx[:, e * H:(e + 1) * H] = d
    return 0


def go2_grouped_linear_wgrad(dY, lddy, X, ldx, dW, M, E, D, H, ws, wsn, stream):
    _need(dY and X and dW and ws and M > 0 and E > 0 and D > 0 and H > 0, "go2_grouped_linear_wgrad: bad argument")
    _need(wsn >= ((M + 255) // 256) * E * D * H, "go2_grouped_linear_wgrad: workspace too small (go2_grouped_linear_wgrad_workspace)")
    g, x, dw = _mat(dY, M, E * D, lddy), _mat(X, M, E * H, ldx), _mat(dW, E * D, H, H)
    for e in range(E):
        dw[e * D:(e + 1) * D] = (g[:, e * D:(e + 1) * D].T @ x[:, e * H:(e + 1) * H]).astype(F)
    return 0


def go2_moe_combine_forward(logits, expert_out, gates, pre, n, E, D, stream):
    _need(E <= 16, "go2_moe_combine_forward: at most 16 experts")
    lg = _mat(logits, n, E, E)
    g = np.exp(lg - lg.max(1, keepdims=True)).astype(F)
    g = (g / g.sum(1, keepdims=True, dtype=F)).astype(F)
    _mat(gates, n, E, E)[...] = g
    eo = _mat(expert_out, n, E * D, E * D).reshape(n, E, D)
    _mat(pre, n, D, D)[...] = (g[:, :, None] * eo).sum(1, dtype=F)
    return 0


def go2_gate_usage(gates, usage, n, E, scale, stream):
    _need(E <= 16, "go2_gate_usage: at most 16 experts")
    _vec(usage, E)[...] = _mat(gates, n, E, E).sum(0, dtype=F) / F(n) * F(scale)
    return 0


def go2_moe_combine_backward(dpre, gates, expert_out, usage, lb_coef, dexpert_out, dexpert_out_t, dlogits, dlogits_t, n, E, D, stream):
    _need(E <= 16, "go2_moe_combine_backward: at most 16 experts")
    _vec(usage, E)[...] = _mat(gates, n, E, E).sum(0, dtype=F) / F(n)
    return go2_moe_combine_backward_given_usage(dpre, gates, expert_out, usage, lb_coef, dexpert_out, dexpert_out_t, dlogits, dlogits_t, n, E, D, stream)


def go2_moe_combine_backward_given_usage(dpre, gates, expert_out, usage, lb_coef, dexpert_out, dexpert_out_t, dlogits, dlogits_t, n, E, D, stream):
    g, eo, dp = _mat(gates, n, E, E), _mat(expert_out, n, E * D, E * D).reshape(n, E, D), _mat(dpre, n, D, D)
    us = _vec(usage, E)
    dg = F(lb_coef) * (F(2) / F(E)) * (us - F(1) / F(E)) / F(n) + (dp[:, None, :] * eo).sum(2, dtype=F)
    deo = (g[:, :, None] * dp[:, None, :]).astype(F).reshape(n, E * D)
    _mat(dexpert_out, n, E * D, E * D)[...] = deo
    if dexpert_out_t:
        _mat(dexpert_out_t, E * D, n, n)[...] = deo.T
    dl = (g * (dg - (g * dg).sum(1, keepdims=True, dtype=F))).astype(F)
    _mat(dlogits, n, E, E)[...] = dl
    if dlogits_t:
        _mat(dlogits_t, E, n, n)[...] = dl.T
    return 0


def go2_latent_loss(student, teacher, dstudent, acc, n, d, stream):
    df = _vec(student, n * d) - _vec(teacher, n * d)
    _vec(acc, 1)[0] = (df * df).sum(dtype=F)
    _vec(dstudent, n * d)[...] = F(2) * df / F(n * d)
    return 0


def go2_cts_log(acc, usage, log, count, E, stream):
    lg = _vec(log, 2)
    lg[0] += _vec(acc, 1)[0] / F(count)
    if usage:
        dlt = _vec(usage, E) - F(1) / F(E)
        lg[1] += (dlt * dlt).sum(dtype=F) / F(E)
    return 0


def go2_history_update(history, obs, dones, n, H, d, stream):
    h, o = _mat(history, n, H * d, H * d).reshape(n, H, d), _mat(obs, n, d, d)
    if dones:
        h[_vec(dones, n, C.c_uint8) > 0] = 0
    h[:, :-1] = h[:, 1:].copy()
    h[:, -1] = o
    return 0


_mcp = None


def _kernel_source_entry(name):
    """Entry points whose row arithmetic is the KERNEL'S OWN SOURCE compiled with g++ (tests/emu/emu_mcp.cpp over csrc/mcp_core.cuh)."""
    global _mcp
    if _mcp is None:
        import os
        import subprocess
        d = os.path.join(os.path.dirname(os.path.abspath(__file__)), "emu")
        subprocess.check_call(["make", "-C", d, "-s"])
        _mcp = C.CDLL(os.path.join(d, "libgo2emu_mcp.so"))
    fn = getattr(_mcp, name, None)
    if fn is None:
        return None
    from go2_rl_gym_b200.rl._ops import _SIGS
    fn.argtypes, fn.restype = _SIGS[name], C.c_int

    def checked(*args):
        if fn(*args) != 0:
            raise EmuError(f"{name}: argument check failed")
        return 0
    return checked


class _FakeLib:
    """Attribute access like a ctypes CDLL; unknown entry points fail loudly."""

    launches = 0

    def __getattr__(self, name):
        fn = globals().get(name)
        if fn is None and name.startswith("go2_"):
            fn = _kernel_source_entry(name)
        if fn is None or not name.startswith("go2_"):
            raise AttributeError(f"emu_rl: {name} is not emulated")

        def wrapped(*args):
            _FakeLib.launches += 1
            args = [a.value if isinstance(a, C.c_void_p) else a for a in args]
            try:
                return fn(*[0 if a is None else a for a in args])
            except EmuError as e:
                self._err = str(e)
                return 1
        return wrapped

    def go2_last_error(self):
        return getattr(self, "_err", "").encode()

    def go2_kernel_launch_count(self):
        return _FakeLib.launches


def install(monkeypatch):
    """Route go2_rl_gym_b200.rl._ops through the emulation for the duration of one test (CPU tensors, no CUDA graphs)."""
    from go2_rl_gym_b200.rl import _ops
    fake = _FakeLib()
    monkeypatch.setattr(_ops, "lib", lambda: fake)
    monkeypatch.setattr(_ops, "_stream", lambda: 0)
    monkeypatch.setenv("GO2_GRAPH", "0")
    return fake


def install_process():
    """install() for a process that is not under pytest's monkeypatch (spawned gloo workers)."""
    import os
    from go2_rl_gym_b200.rl import _ops
    fake = _FakeLib()
    _ops.lib = lambda: fake
    _ops._stream = lambda: 0
    os.environ["GO2_GRAPH"] = "0"
    return fake

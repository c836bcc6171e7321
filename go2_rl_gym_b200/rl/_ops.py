"""ctypes bindings of the trainer kernels (include/go2_b200.h, RL section) + a small MLP engine over a flat parameter
vector.  Every call launches hand-written CUDA from libgo2b200.so on torch's current stream; there is no fallback."""
import ctypes as C
import os

import torch

from .. import _abi

_vp, _i, _f, _l = C.c_void_p, C.c_int, C.c_float, C.c_long
_SIGS = {
    "go2_linear_forward_simt": [_vp, _i, _vp, _i, _vp, _vp, _i, _vp, _i, _i, _i, _i, _i, _vp],
    "go2_linear_forward_tc": [_vp, _i, _vp, _i, _vp, _vp, _i, _vp, _i, _i, _i, _i, _i, _vp],
    "go2_linear_dgrad_simt": [_vp, _i, _vp, _i, _vp, _i, _vp, _i, _vp, _i, _i, _i, _i, _vp],
    "go2_linear_dgrad_tc": [_vp, _i, _vp, _i, _vp, _i, _vp, _i, _vp, _i, _vp, _i, _i, _i, _i, _vp],
    "go2_linear_wgrad_simt": [_vp, _i, _vp, _i, _vp, _i, _vp, _i, _i, _i, _vp, _l, _vp],
    "go2_linear_wgrad_tc": [_vp, _i, _vp, _i, _vp, _i, _vp, _i, _i, _i, _vp, _l, _vp],
    "go2_linear_wgrad_tc_rm": [_vp, _i, _vp, _i, _vp, _i, _vp, _i, _i, _i, _vp, _l, _vp],
    "go2_linear_forward_smalln": [_vp, _i, _vp, _i, _vp, _vp, _i, _i, _i, _i, _vp],
    "go2_linear_dgrad_smalln": [_vp, _i, _vp, _i, _vp, _i, _vp, _i, _i, _i, _i, _vp],
    "go2_linear_wgrad_smalln": [_vp, _i, _vp, _i, _vp, _i, _vp, _i, _i, _i, _vp, _l, _vp],
    "go2_linear_wgrad_rank1": [_vp, _i, _vp, _i, _vp, _vp, _i, _i, _vp, _l, _vp],
    "go2_refresh_weights": [_i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp],
    "go2_transpose": [_vp, _i, _vp, _i, _i, _i, _vp],
    "go2_colsum": [_vp, _i, _vp, _i, _i, _vp, _vp],
    "go2_sample_actions": [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, C.c_uint64, C.c_uint32, _i, _vp],
    "go2_sample_actions_dev": [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, C.c_uint64, _vp, _i, _vp],
    "go2_process_env_step": [_vp, _vp, _vp, _vp, _vp, _vp, _i, _f, _vp, _vp],
    "go2_gae": [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _f, _f, _vp, _vp],
    "go2_adv_normalize": [_vp, _l, _vp, C.c_double, _vp],
    "go2_gather_rows": [_vp, _i, _vp, _vp, _i, _vp, _l, _vp],
    "go2_ppo_loss": [_vp] * 14 + [_i, _i, _f, _f, _f, _i, _f, _i, _f, _f, _vp],
    "go2_kl_adaptive_lr": [_vp, _f, _f, _vp, _vp, _f, _f, _vp],
    "go2_concat2": [_vp, _i, _i, _vp, _i, _i, _vp, _i, _vp, _l, _vp],
    "go2_l2norm_forward": [_vp, _i, _vp, _i, _vp, _l, _i, _vp],
    "go2_l2norm_backward": [_vp, _i, _vp, _i, _vp, _vp, _i, _vp, _l, _i, _vp],
    "go2_grouped_linear_forward": [_vp, _l, _vp, _vp, _vp, _l, _l, _i, _i, _i, _vp],
    "go2_grouped_linear_dgrad": [_vp, _l, _vp, _vp, _l, _vp, _l, _l, _i, _i, _i, _vp],
    "go2_grouped_linear_wgrad": [_vp, _l, _vp, _l, _vp, _l, _i, _i, _i, _vp, _l, _vp],
    "go2_moe_combine_forward": [_vp, _vp, _vp, _vp, _l, _i, _i, _vp],
    "go2_moe_combine_backward": [_vp, _vp, _vp, _vp, _f, _vp, _vp, _vp, _vp, _l, _i, _i, _vp],
    "go2_gate_usage": [_vp, _vp, _l, _i, _f, _vp],
    "go2_moe_combine_backward_given_usage": [_vp, _vp, _vp, _vp, _f, _vp, _vp, _vp, _vp, _l, _i, _i, _vp],
    "go2_latent_loss": [_vp, _vp, _vp, _vp, _l, _i, _vp],
    "go2_cts_log": [_vp, _vp, _vp, _l, _i, _vp],
    "go2_history_update": [_vp, _vp, _vp, _l, _i, _i, _vp],
    "go2_gather_u8": [_vp, _vp, _vp, _l, _vp],
    "go2_adam_clip_step": [_vp, _vp, _vp, _vp, _l, _f, _vp, _f, _vp, _vp],
    "go2_mcp_compose_forward": [_vp, _vp, _vp, _vp, _vp, _l, _i, _i, _vp],
    "go2_mcp_compose_backward": [_vp, _vp, _vp, _vp, _vp, _vp, _l, _i, _i, _vp],
    "go2_sample_actions_sigma": [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, C.c_uint64, C.c_uint32, _vp, _i, _vp],
    "go2_allreduce_p2p": [_vp, _vp, _vp, _l, _l, _i, _i, _vp, _vp],
    "go2_allreduce_p2p2": [_vp, _vp, _vp, _vp, _l, _l, _i, _i, _vp, _vp],
    "go2_ppo_loss_sigma": [_vp] * 14 + [_i, _i, _f, _f, _f, _i, _f, _i, _f, _f, _vp],
}
_lib = None


def lib():
    global _lib
    if _lib is None:
        L = _abi.load_library()
        for name, sig in _SIGS.items():
            fn = getattr(L, name)
            fn.argtypes = sig
            fn.restype = C.c_int
        _lib = L
    return _lib


def _stream():
    return torch.cuda.current_stream().cuda_stream


def call(name, *args):
    L = lib()
    rc = getattr(L, name)(*args, _stream())
    if rc != 0:
        raise RuntimeError(f"{name} failed ({rc}): {L.go2_last_error().decode()}")


def ptr(t):
    return 0 if t is None else t.data_ptr()


def upload_steps(buf, first, T, device):
    """int32 device array {first + 1, ..., first + T}: per-step Philox counters of a graph-replayed rollout."""
    host = torch.arange(first + 1, first + T + 1, dtype=torch.int64).to(torch.int32)     # wraps like the uint32 kernel argument
    if buf is None or buf.numel() != T:
        buf = torch.zeros(T, dtype=torch.int32, device=device)
    buf.copy_(host)
    return buf


def use_tc():
    """GO2_GEMM=simt forces the strict-fp32 CUDA-core GEMM everywhere (numerical cross-check of the tensor-core path)."""
    return os.environ.get("GO2_GEMM", "tc") != "simt"


class GraphSet:
    """CUDA-graph cache for launch sequences that depend only on device-resident state.  run(key, fn): first call eager
    (allocations, cudaFuncSetAttribute, tensor-map encodes), second call captured, replayed afterwards.  GO2_GRAPH=0 disables
    graphs; a failed capture makes that key eager for good.  Collectives are never captured: with world_size > 1 the callers cut
    the optimiser step into [gradient graph] -> eager NCCL all-reduce -> [Adam graph]."""

    replayed_launches = 0     # kernel launches issued through graph replays (the library's own counter only sees direct launches)

    def __init__(self):
        self._g, self._warm, self._failed, self._n = {}, set(), set(), {}
        self.enabled = os.environ.get("GO2_GRAPH", "1") != "0"

    def run(self, key, fn):
        if not self.enabled or key in self._failed:
            return fn()
        g = self._g.get(key)
        if g is not None:
            GraphSet.replayed_launches += self._n[key]
            return g.replay()
        if key not in self._warm:
            self._warm.add(key)
            return fn()
        import warnings
        g = torch.cuda.CUDAGraph()
        torch.cuda.synchronize()
        n0 = lib().go2_kernel_launch_count()
        try:
            with torch.cuda.graph(g, capture_error_mode="thread_local"):
                fn()
        except Exception as e:  # noqa: BLE001 - anything the capture rejects
            self._failed.add(key)
            warnings.warn(f"CUDA-graph capture of {key} failed ({type(e).__name__}: {e}); using eager launches")
            torch.cuda.synchronize()
            return fn()
        self._g[key], self._n[key] = g, lib().go2_kernel_launch_count() - n0     # launches counted during capture did not run
        GraphSet.replayed_launches += self._n[key]
        g.replay()


class SideStream:
    """Second CUDA stream for work that is independent of the main chain (the critic next to the actor): fork() makes it wait for everything issued
    so far on the current stream, `with side:` issues on it, join() makes the current stream wait for it.  The GEMM kernels are persistent
    one-CTA-per-SM grids whose tails leave SMs idle (static tile lists, 1.3 - 2.6 waves of tiles): two independent chains fill each other's tails.
    Inside a CUDA-graph capture the events become graph edges.  Disabled (everything on the current stream) on the CPU and with GO2_TWO_STREAMS=0."""

    def __init__(self, device):
        self.enabled = torch.device(device).type == "cuda" and os.environ.get("GO2_TWO_STREAMS", "1") != "0"
        if self.enabled:
            self.stream = torch.cuda.Stream(device=device)
            self._e0, self._e1 = torch.cuda.Event(), torch.cuda.Event()
        self._ctx = None

    def fork(self):
        if self.enabled:
            self._e0.record()
            self.stream.wait_event(self._e0)

    def join(self):
        if self.enabled:
            self._e1.record(self.stream)
            torch.cuda.current_stream().wait_event(self._e1)

    def __enter__(self):
        if self.enabled:
            self._ctx = torch.cuda.stream(self.stream)
            self._ctx.__enter__()
        return self

    def __exit__(self, *exc):
        if self.enabled:
            ctx, self._ctx = self._ctx, None
            return ctx.__exit__(*exc)
        return False


class StreamPool:
    """n side streams for loops of independent small launches (the E block-diagonal experts of a mixture layer: each is a GEMM of a few dozen tiles that
    cannot fill 148 SMs on its own).  fork() -> `with pool.lane(i):` ... -> join().  Lane i is a SideStream; all lanes fork from / join into the
    current stream."""

    def __init__(self, device, n=4):
        self.lanes = [SideStream(device) for _ in range(n)]
        self.n = n
        self.enabled = self.lanes[0].enabled

    def fork(self):
        for l in self.lanes:
            l.fork()

    def lane(self, i):
        return self.lanes[i % self.n]

    def join(self):
        for l in self.lanes:
            l.join()


def _pad4(n):
    return (n + 3) // 4 * 4


def pad_in(width):
    """Row pitch of a tensor-core MLP input of `width` features: room for the ones column, multiple of 4 floats (16 bytes)."""
    return _pad4(width + 1)


class MlpEngine:
    """A chain of Linear(+ELU) layers on views of a flat parameter vector, evaluated by the library's GEMM kernels.

    dims = [in, h1, ..., out]; ELU after every layer but the last (actor_critic.py:58-79), or after the last one too when
    last_act (the experts' backbone, modules/utils.py:81).  Tensor-core path (tcgen05, tf32 multiply / fp32 accumulate):
      * forward / dgrad are K-major contractions: X [M, in] and W [out, in] (dgrad: W^T, refreshed with the weights) need
        16-byte row pitches -> the first layer uses a zero-padded copy of its weight (refreshed after each optimiser step) and
        inputs whose leading dimension is not a multiple of 4 go through a padded staging copy;
      * wgrad contracts over the batch rows and reads the ROW-MAJOR activations / gradients as MN-major operands
        (go2_linear_wgrad_tc_rm): no transposed copies anywhere;
      * every hidden activation buffer is [rows, d + 4] with column d = 1, and padded inputs carry a 1 in their first padding
        column (x_ones), so the bias gradient is one more column of the same wgrad GEMM.
    Layers whose shapes cannot meet the TMA alignment rules (the critic's 1-wide head) run on CUDA-core kernels."""

    def __init__(self, dims, weights, biases, gweights, gbiases, max_rows, device, train_rows=0, last_act=False, need_dx=False):
        self.dims, self.L = list(dims), len(dims) - 1
        self.W, self.b, self.gW, self.gb = weights, biases, gweights, gbiases
        self.max_rows, self.train_rows, self.last_act, self.need_dx = max_rows, train_rows, last_act, need_dx
        self.tc = use_tc()
        dev = device
        hidden = dims[1:] if last_act else dims[1:-1]
        self.acts = [torch.ones(max_rows, d + 4, device=dev) for d in hidden]       # column d stays 1
        hmax = max(dims[1:-1]) if self.L > 1 else dims[-1]
        # one gradient buffer per layer boundary: the weight gradient of layer l (side stream) may still read dZ_l while the chain moves on
        self.dbuf = [torch.empty(max(train_rows, 1), hmax, device=dev) for _ in range(max(self.L - 1, 1) if train_rows else 1)]
        self.wside = SideStream(dev if train_rows else "cpu")        # weight gradients run next to the dgrad chain (disabled for inference-only engines)
        self.work = torch.empty(64 * max(_pad4(dims[l] + 1) * ((dims[l + 1] + 127) // 128 * 128) for l in range(self.L)), device=dev)
        self.kpad0 = pad_in(dims[0]) if self.tc else dims[0]               # room for the ones column
        self.W0p = torch.zeros(dims[1], self.kpad0, device=dev) if self.tc else None
        first_wt = 0 if need_dx else 1
        self.Wt = ([None] * first_wt + [torch.zeros(dims[l], _pad4(dims[l + 1]), device=dev) for l in range(first_wt, self.L)]) if self.tc else None
        self._xpad = torch.zeros(max_rows, self.kpad0, device=dev) if self.tc else None
        self.dx = torch.zeros(max(train_rows, 1), self.kpad0, device=dev) if need_dx else None
        self._dirty_w0, self._dirty_wt = True, True
        self._jobs = None

    def mark_dirty(self):
        self._dirty_w0, self._dirty_wt = True, True

    def _refresh(self, need_wt):
        """Operand copies derived from the weights, rebuilt in ONE launch after the weights changed: the zero-padded first-layer weight and
        W^T of every layer whose dgrad runs on the tensor cores (narrow heads read W directly)."""
        if not self.tc or not (self._dirty_w0 or (need_wt and self._dirty_wt)):
            return
        if self._jobs is None:
            jobs = [(self.W[0], self.dims[0], self.W0p, self.kpad0, self.dims[1], self.dims[0], 0)]
            for l in range(0 if self.need_dx else 1, self.L):
                if self.dims[l + 1] <= 16 and self.dims[l] <= 128 and l > 0:
                    continue
                jobs.append((self.W[l], self.dims[l], self.Wt[l], self.Wt[l].shape[1], self.dims[l + 1], self.dims[l], 1))
            assert len(jobs) <= 8
            n = len(jobs)
            vp, ia = (C.c_void_p * n), (C.c_int * n)
            self._jobs = (n, vp(*[j[0].data_ptr() for j in jobs]), ia(*[j[1] for j in jobs]), vp(*[j[2].data_ptr() for j in jobs]),
                          ia(*[j[3] for j in jobs]), ia(*[j[4] for j in jobs]), ia(*[j[5] for j in jobs]), ia(*[j[6] for j in jobs]))
        call("go2_refresh_weights", *self._jobs)
        self._dirty_w0 = self._dirty_wt = False

    @property
    def out(self):
        """Last-layer activation buffer (last_act engines): [rows, dims[-1] + 4], leading dimension ld_out."""
        return self.acts[-1]

    @property
    def ld_out(self):
        return self.acts[-1].shape[1]

    def forward(self, X, ldx, M, out=None, ld_out=0, train=False, x_ones=False, **_unused):
        """out[M, dims[-1]] = MLP(X[M, dims[0]]).  train=True keeps what backward() needs.  x_ones: X[:, dims[0]] == 1 (ldx > dims[0]),
        e.g. rows written by go2_gather_rows / go2_concat2 with a padded pitch.  last_act engines write their output into self.out."""
        assert M <= self.max_rows and (not train or M <= self.train_rows)
        self._refresh(need_wt=train)
        src, lds = X, ldx
        if self.tc and (ldx % 4 or X.data_ptr() % 16 or ldx < self.kpad0):
            # inputs straight from the env rows (ld 45 / 263): padded staging copy (zeros, ones column at dims[0])
            call("go2_gather_rows", ptr(X), self.dims[0], 0, ptr(self._xpad), self.kpad0, 0, M)
            src, lds, x_ones = self._xpad, self.kpad0, True
        self._Xin, self._ldxin, self._x_ones = src, lds, x_ones
        for l in range(self.L):
            last = l == self.L - 1
            if last and not self.last_act:
                dst, ldd = out, ld_out
            else:
                dst, ldd = self.acts[l], self.dims[l + 1] + 4
            act = 0 if (last and not self.last_act) else 1
            if act == 0 and self.dims[l + 1] <= 16 and self.dims[l] <= 128 and l > 0:
                # narrow heads (12 actions, 1 value): streaming fp32 kernel instead of a 128-wide tensor-core tile
                call("go2_linear_forward_smalln", ptr(src), lds, ptr(self.W[l]), self.dims[l], ptr(self.b[l]), ptr(dst), ldd, M, self.dims[l + 1], self.dims[l])
            elif self.tc:
                # layer 0 contracts over the padded width: the padding columns of W0p are zero
                W, ldw, K = (self.W0p, self.kpad0, self.kpad0) if l == 0 else (self.W[l], self.dims[l], self.dims[l])
                call("go2_linear_forward_tc", ptr(src), lds, ptr(W), ldw, ptr(self.b[l]), ptr(dst), ldd, 0, 0, M, self.dims[l + 1], K, act)
            else:
                call("go2_linear_forward_simt", ptr(src), lds, ptr(self.W[l]), self.dims[l], ptr(self.b[l]), ptr(dst), ldd, 0, 0, M,
                     self.dims[l + 1], self.dims[l], act)
            src, lds = dst, ldd
        self._M = M

    def backward(self, dY, lddy, *_unused):
        """Overwrites gW/gb with d loss / d params for the rows of the last forward(train=True).
        dY [M, out] = gradient w.r.t. the LAST LINEAR's output (for last_act engines the caller has already applied ELU').
        With need_dx the gradient w.r.t. the input lands in self.dx [M, kpad0]."""
        M = self._M
        d, ldd = dY, lddy
        ws = self.wside
        for l in range(self.L - 1, -1, -1):
            n_out, n_in = self.dims[l + 1], self.dims[l]
            xin, ldx = (self._Xin, self._ldxin) if l == 0 else (self.acts[l - 1], n_in + 4)
            ones = self._x_ones if l == 0 else True
            small = n_out <= 16 and n_in <= 128 and l > 0
            # weight gradient of layer l: needs dZ_l (ready on the current stream) and the stored activations; nothing of the dgrad chain needs it.
            # All weight gradients of this engine share `work`, so they stay in order on ONE side stream.
            ws.fork()
            with ws:
                if small:       # narrow heads: streaming fp32 kernels (weights + bias in one pass)
                    call("go2_linear_wgrad_smalln", ptr(d), ldd, ptr(xin), ldx, ptr(self.gW[l]), n_in, ptr(self.gb[l]), M, n_out, n_in, ptr(self.work), self.work.numel())
                elif self.tc and ldd % 4 == 0 and d.data_ptr() % 16 == 0:
                    if not ones:
                        call("go2_colsum", ptr(d), ldd, ptr(self.gb[l]), M, n_out, ptr(self.work))
                    call("go2_linear_wgrad_tc_rm", ptr(d), ldd, ptr(xin), ldx, ptr(self.gW[l]), n_in, ptr(self.gb[l]) if ones else 0, M, n_out, n_in,
                         ptr(self.work), self.work.numel())
                elif n_out == 1 and n_in <= 512:      # the critic's scalar head: one streaming pass (weights + bias)
                    call("go2_linear_wgrad_rank1", ptr(d), ldd, ptr(xin), ldx, ptr(self.gW[l]), ptr(self.gb[l]), M, n_in, ptr(self.work), self.work.numel())
                else:
                    call("go2_colsum", ptr(d), ldd, ptr(self.gb[l]), M, n_out, ptr(self.work))
                    call("go2_linear_wgrad_simt", ptr(d), ldd, ptr(xin), ldx, ptr(self.gW[l]), n_in, 0, M, n_out, n_in, ptr(self.work), self.work.numel())
            if l > 0:
                nxt = self.dbuf[l - 1]
                if small:
                    call("go2_linear_dgrad_smalln", ptr(d), ldd, ptr(self.W[l]), n_in, ptr(self.acts[l - 1]), n_in + 4, ptr(nxt), n_in, M, n_out, n_in)
                elif self.tc and n_out % 4 == 0:
                    call("go2_linear_dgrad_tc", ptr(d), ldd, ptr(self.Wt[l]), self.Wt[l].shape[1], ptr(self.acts[l - 1]), n_in + 4, 0, 0,
                         ptr(nxt), n_in, 0, 0, M, n_out, n_in)
                else:
                    call("go2_linear_dgrad_simt", ptr(d), ldd, ptr(self.W[l]), n_in, ptr(self.acts[l - 1]), n_in + 4, ptr(nxt), n_in, 0, 0, M, n_out, n_in)
                d, ldd = nxt, n_in
            elif self.need_dx:
                if self.tc and n_out % 4 == 0:
                    call("go2_linear_dgrad_tc", ptr(d), ldd, ptr(self.Wt[0]), self.Wt[0].shape[1], 0, 0, 0, 0, ptr(self.dx), self.kpad0, 0, 0, M, n_out, n_in)
                else:
                    call("go2_linear_dgrad_simt", ptr(d), ldd, ptr(self.W[0]), n_in, 0, 0, ptr(self.dx), self.kpad0, 0, 0, M, n_out, n_in)
        ws.join()

"""ctypes bindings of the trainer kernels (include/go2_b200.h, RL section) + a small MLP engine over a flat parameter
vector.  Every call launches hand-written CUDA from libgo2b200.so on torch's current stream; there is no fallback."""
import ctypes as C
import os

import torch

from .. import _abi

_vp, _i, _f, _l = C.c_void_p, C.c_int, C.c_float, C.c_long
_SIGS = {
    "go2_linear_forward_simt": [_vp, _i, _vp, _i, _vp, _vp, _i, _i, _i, _i, _i, _vp],
    "go2_linear_dgrad_simt": [_vp, _i, _vp, _i, _vp, _i, _vp, _i, _i, _i, _i, _vp],
    "go2_linear_wgrad_simt": [_vp, _i, _vp, _i, _vp, _i, _vp, _i, _i, _i, _vp, _l, _vp],
    "go2_sample_actions": [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, C.c_uint64, C.c_uint32, _i, _vp],
    "go2_process_env_step": [_vp, _vp, _vp, _vp, _vp, _vp, _i, _f, _vp],
    "go2_gae": [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _f, _f, _vp, _vp],
    "go2_adv_normalize": [_vp, _l, _vp, C.c_double, _vp],
    "go2_gather_rows": [_vp, _i, _vp, _vp, _i, _l, _vp],
    "go2_ppo_loss": [_vp] * 13 + [_i, _i, _f, _f, _f, _i, _f, _vp],
    "go2_kl_adaptive_lr": [_vp, _f, _f, _vp, _vp, _vp],
    "go2_adam_clip_step": [_vp, _vp, _vp, _vp, _l, _f, _vp, _i, _f, _vp, _vp],
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


USE_TC = os.environ.get("GO2_GEMM", "tc") != "simt"


class MlpEngine:
    """A chain of Linear(+ELU) layers evaluated with the library's GEMM kernels on views of a flat parameter vector.

    dims = [in, h1, ..., out]; ELU after every layer but the last (actor_critic.py:58-79).  `weights[l]`/`biases[l]` are
    views into the flat parameter buffer, `gweights[l]`/`gbiases[l]` the matching views into the flat gradient buffer."""

    def __init__(self, dims, weights, biases, gweights, gbiases, max_rows, device):
        self.dims, self.L = list(dims), len(dims) - 1
        self.W, self.b, self.gW, self.gb = weights, biases, gweights, gbiases
        self.max_rows = max_rows
        self.acts = [torch.empty(max_rows, d, device=device) for d in dims[1:-1]]
        hmax = max(dims[1:-1]) if self.L > 1 else dims[-1]
        self.dbuf = [torch.empty(max_rows, hmax, device=device) for _ in range(2)]
        self.work = torch.empty(64 * max(dims[l] * dims[l + 1] for l in range(self.L)), device=device)

    def forward(self, X, ldx, M, out, ld_out, save=True):
        """out[M, dims[-1]] = MLP(X[M, dims[0]]); keeps the hidden activations when save (needed by backward)."""
        assert M <= self.max_rows
        src, lds = X, ldx
        for l in range(self.L):
            last = l == self.L - 1
            dst, ldd = (out, ld_out) if last else (self.acts[l], self.dims[l + 1])
            call("go2_linear_forward_simt", ptr(src), lds, ptr(self.W[l]), self.dims[l], ptr(self.b[l]), ptr(dst), ldd, M, self.dims[l + 1],
                 self.dims[l], 0 if last else 1)
            src, lds = dst, ldd
        self._X, self._ldx, self._M = X, ldx, M

    def backward(self, dY, lddy):
        """Accumulates nothing: overwrites gW/gb with d loss / d params for the rows of the last forward()."""
        M = self._M
        d, ldd = dY, lddy
        for l in range(self.L - 1, -1, -1):
            xin, ldx = (self._X, self._ldx) if l == 0 else (self.acts[l - 1], self.dims[l])
            call("go2_linear_wgrad_simt", ptr(d), ldd, ptr(xin), ldx, ptr(self.gW[l]), self.dims[l], ptr(self.gb[l]), M, self.dims[l + 1],
                 self.dims[l], ptr(self.work), self.work.numel())
            if l > 0:
                nxt = self.dbuf[l % 2]
                call("go2_linear_dgrad_simt", ptr(d), ldd, ptr(self.W[l]), self.dims[l], ptr(self.acts[l - 1]), self.dims[l], ptr(nxt),
                     self.dims[l], M, self.dims[l + 1], self.dims[l])
                d, ldd = nxt, self.dims[l]

"""Host-side dispatcher: the part of the reference's C wrappers that is not arithmetic.

It reproduces, for numpy inputs and for device-resident torch CUDA tensors, what the
reference does around its per-element loops (hpgeom/hpgeom.c:183-238 and the sister functions):
coercion to int64/float64 with *safe* casting (PyArray_FROM_OTF), broadcasting nside against
the data (NpyIter_MultiNew), the fixed broadcast-failure messages, empty inputs, 0-d results
returned as numpy scalars (PyArray_Return), and ValueError with the reference's message for
the FIRST offending element.  All arithmetic happens in libhpgb.so's CUDA kernels.

numpy in  -> hpgb_host_<fn>() (chunked H2D / kernel / D2H pipeline)      -> numpy out
torch.cuda in -> hpgb_<fn>() on the tensor's device and current stream   -> torch.cuda out
"""
import ctypes

import numpy as np

from . import _lib

try:  # torch is plumbing (device memory + streams); numpy-only use works without it
    import torch
except Exception:  # pragma: no cover
    torch = None

I8 = "i8"
F8 = "f8"
_NP = {I8: np.int64, F8: np.float64}


def _is_torch(x):
    return torch is not None and isinstance(x, torch.Tensor)


# ----------------------------------------------------------------------------------------------
# numpy backend
# ----------------------------------------------------------------------------------------------
class _NumpyBackend:
    name = "numpy"

    def __init__(self, device=0):
        self.device = device

    @staticmethod
    def coerce(obj, kind):
        arr = np.asarray(obj)
        want = _NP[kind]
        if arr.dtype != want:
            # same rule (and TypeError text) as PyArray_FROM_OTF in the reference
            arr = arr.astype(want, casting="safe")
        return arr

    @staticmethod
    def shape(a):
        return tuple(a.shape)

    @staticmethod
    def expand(a, shape):
        if tuple(a.shape) != tuple(shape):
            a = np.broadcast_to(a, shape)
        return np.ascontiguousarray(a)

    @staticmethod
    def empty(shape, kind):
        return np.empty(shape, dtype=_NP[kind])

    @staticmethod
    def ptr(a):
        return ctypes.c_void_p(a.ctypes.data) if a is not None else None

    @staticmethod
    def item(a, i):
        return a.reshape(-1)[i].item()

    @staticmethod
    def scalar(a):
        return a[()]  # 0-d array -> numpy scalar, like PyArray_Return

    def call(self, fn, ptr_args, tail_args):
        """fn = C name without the hpgb_ prefix; returns first_bad (-1 if none)."""
        first_bad = ctypes.c_int64(-1)
        f = getattr(_lib.lib(), "hpgb_host_" + fn)
        rc = f(*ptr_args, *tail_args, self.device, ctypes.byref(first_bad))
        return rc, first_bad.value


# ----------------------------------------------------------------------------------------------
# torch (device resident) backend
# ----------------------------------------------------------------------------------------------
class _TorchBackend:
    name = "torch"
    _status = {}

    def __init__(self, device):
        self.dev = device
        self.device = device.index if device.index is not None else torch.cuda.current_device()

    def coerce(self, obj, kind):
        want = torch.int64 if kind == I8 else torch.float64
        if _is_torch(obj):
            t = obj
            if t.device.type != "cuda":
                t = t.to(self.dev)
            elif t.device != self.dev and t.device.index != self.device:
                raise ValueError("all tensors must live on the same CUDA device")
            if t.dtype != want:
                if not torch.can_cast(t.dtype, want) or (kind == I8 and t.dtype.is_floating_point) \
                        or t.dtype == torch.uint64:
                    raise TypeError("Cannot cast array data from dtype('%s') to dtype('%s') according to the "
                                    "rule 'safe'" % (str(t.dtype).replace("torch.", ""), "int64" if kind == I8 else "float64"))
                t = t.to(want)
            return t
        arr = _NumpyBackend.coerce(obj, kind)
        return torch.from_numpy(np.ascontiguousarray(arr)).to(self.dev)

    @staticmethod
    def shape(a):
        return tuple(a.shape)

    @staticmethod
    def expand(a, shape):
        if tuple(a.shape) != tuple(shape):
            a = a.expand(shape)
        return a.contiguous()

    def empty(self, shape, kind):
        return torch.empty(shape, dtype=torch.int64 if kind == I8 else torch.float64, device=self.dev)

    @staticmethod
    def ptr(a):
        return ctypes.c_void_p(a.data_ptr()) if a is not None else None

    @staticmethod
    def item(a, i):
        return a.reshape(-1)[i].item()

    @staticmethod
    def scalar(a):
        return a  # 0-d tensors stay 0-d tensors on the device

    def _status_buf(self):
        st = self._status.get(self.device)
        if st is None:
            st = torch.empty(2, dtype=torch.int64, device=self.dev)
            self._status[self.device] = st
        return st

    def call(self, fn, ptr_args, tail_args):
        L = _lib.lib()
        with torch.cuda.device(self.device):
            st = self._status_buf()
            stream = ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
            stp = ctypes.c_void_p(st.data_ptr())
            rc = L.hpgb_status_reset(stp, stream)
            if rc == 0:
                rc = getattr(L, "hpgb_" + fn)(*ptr_args, *tail_args, stp, stream)
            if rc != 0:
                return rc, -1
            first = int(st[0].item())  # 8-byte D2H; synchronises the stream like the numpy path does
        return 0, first  # first == -1 <=> HPGB_NO_BAD_ELEMENT (all ones)


def pick_backend(*objs):
    """torch backend iff any operand is a CUDA tensor; numpy otherwise."""
    for o in objs:
        if _is_torch(o) and o.device.type == "cuda":
            return _TorchBackend(o.device)
    for o in objs:
        if _is_torch(o):  # CPU tensors are treated like numpy arrays
            break
    return _NumpyBackend(default_device())


_default_device = 0


def default_device():
    return _default_device


def set_default_device(index):
    """GPU used for numpy inputs (torch inputs run where they live)."""
    global _default_device
    _default_device = int(index)


def _np_view(o):
    return o.detach().cpu().numpy() if _is_torch(o) else o


# ----------------------------------------------------------------------------------------------
# the generic elementwise call
# ----------------------------------------------------------------------------------------------
def elementwise(fn, nside, ins, in_kinds, out_kinds, tail, *, nest, bcast_msg, explain,
                out_cols=None, max_ndim=None, keep_0d_row=False):
    """Run one conversion.

    fn        : C entry point without prefix ("angle_to_pixel", ...)
    ins       : data operands (after nside), in the C argument order
    out_kinds : dtype kind of every output; out_cols[k] = trailing row length or None
    tail      : integer flags that follow (n, nside, nside_v) in the C signature
    explain   : callable(nside_i, values_i) -> message for the first bad element
    """
    be = pick_backend(nside, *ins)
    if be.name == "numpy":
        nside, ins = _np_view(nside), [_np_view(x) for x in ins]
    ns = be.coerce(nside, I8)
    arrs = [be.coerce(x, k) for x, k in zip(ins, in_kinds)]
    if max_ndim is not None:
        max_ndim(ns, arrs)
    try:
        shape = np.broadcast_shapes(be.shape(ns), *[be.shape(a) for a in arrs])
    except ValueError:
        raise ValueError(bcast_msg) from None
    out_cols = out_cols or [None] * len(out_kinds)
    n = int(np.prod(shape, dtype=np.int64))
    zero_d = len(shape) == 0

    def out_shape(cols):
        if cols is None:
            return shape
        return (cols,) if zero_d else (n, cols)

    outs = [be.empty(out_shape(c), k) for k, c in zip(out_kinds, out_cols)]
    if n == 0:
        return outs
    scalar_nside = len(be.shape(ns)) == 0 or int(np.prod(be.shape(ns))) == 1
    if scalar_nside:
        nside_val = int(be.item(ns, 0))
        msg = _lib.explain_nside(nside_val, nest)
        if msg:
            raise ValueError(msg)
        ns_full = None
    else:
        nside_val = 0
        ns_full = be.expand(ns, shape)
    full = [be.expand(a, shape) for a in arrs]
    ptrs = [be.ptr(a) for a in full] + [be.ptr(o) for o in outs]
    rc, first_bad = be.call(fn, ptrs, [n, nside_val, be.ptr(ns_full)] + list(tail))
    _lib.check_rc(rc, nside_val, nest)
    if first_bad >= 0:
        ns_i = nside_val if ns_full is None else int(be.item(ns_full, first_bad))
        raise ValueError(explain(ns_i, [be.item(a, first_bad) for a in full]))
    if zero_d:
        outs = [be.scalar(o) if c is None else o for o, c in zip(outs, out_cols)]
    return outs

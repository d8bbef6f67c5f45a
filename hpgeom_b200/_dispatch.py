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
import contextlib
import ctypes
import threading

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

    def call(self, fn, ptr_args, tail_args, on_bad=None):
        """fn = C name without the hpgb_ prefix; returns first_bad (-1 if none).  The host pipeline is
        synchronous (the outputs are numpy arrays), so checks are never deferred here."""
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

    def call(self, fn, ptr_args, tail_args, on_bad=None):
        """Launch on the tensors' device and the current torch stream.  The status word is a fresh 16-byte
        tensor per call (two threads or two streams never share one).  Normally its 8-byte read-back
        synchronises the stream, like the reference's return does; inside ``deferred_checks()`` nothing is
        read: the call returns at once and the check happens when the context's object is asked."""
        L = _lib.lib()
        with torch.cuda.device(self.device):
            st = torch.empty(2, dtype=torch.int64, device=self.dev)
            stream = ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
            stp = ctypes.c_void_p(st.data_ptr())
            rc = L.hpgb_status_reset(stp, stream)
            if rc == 0:
                rc = getattr(L, "hpgb_" + fn)(*ptr_args, *tail_args, stp, stream)
            if rc != 0:
                return rc, -1
            pending = getattr(_tls, "pending", None)
            if pending is not None:
                pending._add(st, on_bad)
                return 0, -1
            first = int(st[0].item())  # 8-byte D2H; synchronises the stream like the numpy path does
        return 0, first  # first == -1 <=> HPGB_NO_BAD_ELEMENT (all ones)


_tls = threading.local()


class PendingChecks:
    """Element checks of the device-resident calls made inside ``deferred_checks()``.

    The kernels never trap: a bad element gets -1 / 0 and its index goes to the call's status word.  This object
    keeps those status words; ``raise_if_bad()`` reads them (ONE synchronisation for the whole chain) and raises
    the reference's ValueError for the first failing call, in call order."""

    def __init__(self):
        self._items = []

    def _add(self, status, on_bad):
        self._items.append((status, on_bad))

    def __len__(self):
        return len(self._items)

    def raise_if_bad(self):
        items, self._items = self._items, []
        if not items:
            return
        firsts = torch.stack([st[0] for st, _ in items]).cpu().tolist()  # one D2H copy
        for first, (_, on_bad) in zip(firsts, items):
            if first >= 0:
                if on_bad is not None:
                    on_bad(int(first))
                raise ValueError("element %d failed the reference's range check" % first)


@contextlib.contextmanager
def deferred_checks():
    """Run device-resident (torch CUDA) calls without the per-call stream synchronisation.

        with hpg.deferred_checks() as pending:
            pix = hpg.angle_to_pixel(nside, lon, lat)      # returns at once, kernel queued
            ring = hpg.nest_to_ring(nside, pix)            # chained on the same stream
        pending.raise_if_bad()                             # one sync; ValueError as the reference would raise

    The reference releases the GIL around its loops for the same reason (hpgeom/hpgeom.c:287,320): the caller's
    thread should not wait per call.  numpy inputs are unaffected (their outputs must be complete on return)."""
    prev = getattr(_tls, "pending", None)
    pending = PendingChecks()
    _tls.pending = pending
    try:
        yield pending
    finally:
        _tls.pending = prev


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
    """GPU used for numpy inputs (torch inputs run where they live): a device ordinal, or "all" = shard every
    numpy call over the GPUs of the box from this one process (contiguous shards, one host thread and one
    H2D / kernel / D2H pipeline per GPU; results land in the returned numpy arrays).  ``set_devices([...])``
    restricts "all" to a subset."""
    global _default_device
    _default_device = -1 if (isinstance(index, str) and index.lower() == "all") else int(index)


def set_devices(devices=None):
    """The GPUs ``set_default_device("all")`` shards over (None = every visible device)."""
    devs = [] if devices is None else [int(d) for d in devices]
    arr = (ctypes.c_int * max(len(devs), 1))(*devs)
    _lib.check_rc(_lib.lib().hpgb_host_set_devices(arr, len(devs)))


def host_device():
    """`device` argument for hpgb_host_* calls: the default ordinal, or HPGB_DEVICE_ALL (-1)."""
    return _default_device


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
    def on_bad(first_bad):
        ns_i = nside_val if ns_full is None else int(be.item(ns_full, first_bad))
        raise ValueError(explain(ns_i, [be.item(a, first_bad) for a in full]))

    rc, first_bad = be.call(fn, ptrs, [n, nside_val, be.ptr(ns_full)] + list(tail), on_bad)
    _lib.check_rc(rc, nside_val, nest)
    if first_bad >= 0:
        on_bad(first_bad)
    if zero_d:
        outs = [be.scalar(o) if c is None else o for o, c in zip(outs, out_cols)]
    return outs

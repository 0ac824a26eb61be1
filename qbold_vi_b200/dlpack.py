"""DLPack hand-off at the Python boundary (BASELINE.json north_star: "ctypes, with DLPack for tensor hand-off").

A TensorFlow (``tf.experimental.dlpack.to_dlpack``), CuPy, JAX or PyTorch array reaches the C-ABI as a raw device
pointer without passing through torch: this module unpacks the ``DLManagedTensor`` of a DLPack capsule (or of any object
with ``__dlpack__``) with ctypes only, checks what the kernels require (CUDA device memory, float32 / int32 / int64,
C-contiguous) and keeps the producer's buffer alive until the view is released.

    view = DLPackView(tf.experimental.dlpack.to_dlpack(x))         # capsule
    view = DLPackView(cupy_array, stream=stream_handle)            # __dlpack__ protocol, producer syncs with `stream`
    forward_dlpack(layer, oef_dbv, signal_out, stream=0)           # qbold_forward on foreign buffers

Stream semantics follow the array-API protocol: ``stream`` is the CONSUMER's CUDA stream handle (1 = legacy default
stream, 2 = per-thread default, otherwise the ``cudaStream_t`` value); the producer makes its pending work visible to
that stream before handing the pointer over.  Capsules carry no stream: the caller orders the work itself.
"""
from __future__ import annotations

import ctypes as C

kDLCPU, kDLCUDA, kDLCUDAHost, kDLCUDAManaged = 1, 2, 3, 13
kDLInt, kDLUInt, kDLFloat = 0, 1, 2


class DLDevice(C.Structure):
    _fields_ = [('device_type', C.c_int32), ('device_id', C.c_int32)]


class DLDataType(C.Structure):
    _fields_ = [('code', C.c_uint8), ('bits', C.c_uint8), ('lanes', C.c_uint16)]


class DLTensor(C.Structure):
    _fields_ = [('data', C.c_void_p), ('device', DLDevice), ('ndim', C.c_int32), ('dtype', DLDataType),
                ('shape', C.POINTER(C.c_int64)), ('strides', C.POINTER(C.c_int64)), ('byte_offset', C.c_uint64)]


class DLManagedTensor(C.Structure):
    pass


_DELETER = C.CFUNCTYPE(None, C.POINTER(DLManagedTensor))
DLManagedTensor._fields_ = [('dl_tensor', DLTensor), ('manager_ctx', C.c_void_p), ('deleter', _DELETER)]

_api = C.pythonapi
_api.PyCapsule_IsValid.restype = C.c_int
_api.PyCapsule_IsValid.argtypes = [C.py_object, C.c_char_p]
_api.PyCapsule_GetPointer.restype = C.c_void_p
_api.PyCapsule_GetPointer.argtypes = [C.py_object, C.c_char_p]
_api.PyCapsule_SetName.restype = C.c_int
_api.PyCapsule_SetName.argtypes = [C.py_object, C.c_char_p]
_NAME, _USED = b'dltensor', b'used_dltensor'


class DLPackError(TypeError):
    pass


class DLPackView:
    """Borrowed view of a DLPack tensor: ``ptr`` (device address incl. byte offset), ``shape``, ``dtype`` ('f4', 'i4',
    'i8'), ``device_id``.  Owns the capsule: ``release()`` (or garbage collection) calls the producer's deleter."""

    def __init__(self, obj, stream=None, dtype='f4', allow_host=False):
        if hasattr(obj, '__dlpack__') and not _is_capsule(obj):
            kw = {}
            dev_type = None
            if hasattr(obj, '__dlpack_device__'):
                dev_type = int(obj.__dlpack_device__()[0])
                if dev_type not in (kDLCUDA, kDLCUDAManaged) and not allow_host:
                    raise DLPackError('qbold_vi_b200 runs on CUDA device memory only (DLPack device type %d); there is no '
                                      'CPU path' % dev_type)
            if stream is not None and dev_type in (None, kDLCUDA, kDLCUDAManaged):
                kw['stream'] = int(stream)
            try:
                capsule = obj.__dlpack__(**kw)
            except TypeError:                                   # producers that predate the stream keyword
                capsule = obj.__dlpack__()
        else:
            capsule = obj
        if not _is_capsule(capsule):
            raise DLPackError('expected a DLPack capsule or an object with __dlpack__, got %r' % type(obj).__name__)
        if not _api.PyCapsule_IsValid(capsule, _NAME):
            raise DLPackError('the DLPack capsule was already consumed (a capsule can be used once)')
        self._capsule = capsule
        self._managed = C.cast(_api.PyCapsule_GetPointer(capsule, _NAME), C.POINTER(DLManagedTensor))
        _api.PyCapsule_SetName(capsule, _USED)                  # we own it now: the capsule destructor must not free it
        t = self._managed.contents.dl_tensor
        try:
            self.device_type, self.device_id = int(t.device.device_type), int(t.device.device_id)
            ok_dev = (kDLCUDA, kDLCUDAManaged) + ((kDLCPU, kDLCUDAHost) if allow_host else ())
            if self.device_type not in ok_dev:
                raise DLPackError('qbold_vi_b200 runs on CUDA device memory only (DLPack device type %d); there is no CPU '
                                  'path' % self.device_type)
            code = {(kDLFloat, 32): 'f4', (kDLInt, 32): 'i4', (kDLInt, 64): 'i8'}.get((int(t.dtype.code), int(t.dtype.bits)))
            if code is None or int(t.dtype.lanes) != 1:
                raise DLPackError('unsupported DLPack dtype (code %d, %d bits, %d lanes)'
                                  % (t.dtype.code, t.dtype.bits, t.dtype.lanes))
            if dtype is not None and code != dtype:
                raise DLPackError('expected a %s tensor, got %s' % (dtype, code))
            self.dtype = code
            self.shape = tuple(int(t.shape[i]) for i in range(t.ndim))
            if t.strides:                                       # NULL strides = compact row-major
                expect, strides = 1, [int(t.strides[i]) for i in range(t.ndim)]
                for dim, st in zip(reversed(self.shape), reversed(strides)):
                    if dim != 1 and st != expect:
                        raise DLPackError('expected a C-contiguous tensor, got strides %s for shape %s'
                                          % (strides, self.shape))
                    expect *= dim
            self.ptr = int(t.data or 0) + int(t.byte_offset)
            n = 1
            for d in self.shape:
                n *= d
            self.numel = n
            if n > 0 and not self.ptr:
                raise DLPackError('DLPack tensor with a NULL data pointer')
        except Exception:
            self.release()
            raise

    def release(self):
        m, self._managed = getattr(self, '_managed', None), None
        if m is not None and m.contents.deleter:
            m.contents.deleter(m)

    __del__ = release

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.release()


def _is_capsule(obj):
    return type(obj).__name__ == 'PyCapsule'


def _layer_call(layer, fn_name, views, args):
    from . import _lib
    with_dev = views[0].device_id
    for v in views[1:]:
        if v is not None and v.device_id != with_dev:
            raise DLPackError('all tensors of one call must live on the same CUDA device')
    _lib.check(getattr(_lib.lib(), fn_name)(C.byref(layer.params), *args))


def forward_dlpack(layer, oef_dbv, signal_out, stream=0):
    """``qbold_forward`` on foreign buffers: ``oef_dbv`` [..., 2|3] and ``signal_out`` [..., n_tau] are DLPack capsules or
    ``__dlpack__`` producers (float32, CUDA, contiguous); the launch goes to the CUDA stream handle ``stream``
    (0 = legacy default) of the CURRENT device.  Returns nothing: the result is in ``signal_out``'s memory."""
    width = 3 if layer._variable_hct else 2
    with DLPackView(oef_dbv, stream or 1) as x, DLPackView(signal_out, stream or 1) as y:
        if not x.shape or x.shape[-1] != width:
            raise AssertionError('Input should have %d elements in last dimension' % width)
        n = x.numel // width
        if y.numel != n * layer.n_tau:
            raise DLPackError('signal_out holds %d values, expected %d voxels x %d taus' % (y.numel, n, layer.n_tau))
        _layer_call(layer, 'qbold_forward', (x, y), (C.c_void_p(x.ptr), width, n, C.c_void_p(y.ptr), C.c_void_p(stream)))


def forward_backward_dlpack(layer, oef_dbv, g_signal, signal_out, grad_out, stream=0):
    """``qbold_forward_backward`` on foreign buffers (``g_signal`` / ``signal_out`` may be None)."""
    if layer._variable_hct:
        raise DLPackError('use the tensor API for variable_hct gradients')
    with DLPackView(oef_dbv, stream or 1) as x, DLPackView(grad_out, stream or 1) as g:
        gs = DLPackView(g_signal, stream or 1) if g_signal is not None else None
        so = DLPackView(signal_out, stream or 1) if signal_out is not None else None
        try:
            n = x.numel // 2
            if x.shape[-1] != 2 or g.numel != 2 * n:
                raise DLPackError('oef_dbv [...,2] and grad_out [...,2] must hold the same voxels')
            for v in (gs, so):
                if v is not None and v.numel != n * layer.n_tau:
                    raise DLPackError('expected %d voxels x %d taus, got %d values' % (n, layer.n_tau, v.numel))
            _layer_call(layer, 'qbold_forward_backward', (x, g, gs, so),
                        (C.c_void_p(x.ptr), C.c_void_p(gs.ptr) if gs else None, n, C.c_void_p(so.ptr) if so else None,
                         C.c_void_p(g.ptr), C.c_void_p(stream)))
        finally:
            for v in (gs, so):
                if v is not None:
                    v.release()

"""Engine lifetime: device selection and the CUDA stream all kernels are enqueued on."""
from . import _ffi


def init(device=0, stream=None):
    """pcs_init: bind the engine to `device`; `stream` = raw cudaStream_t (int) or None."""
    _ffi.check(_ffi.lib().pcs_init(int(device), None if stream is None else int(stream)))


def shutdown():
    _ffi.lib().pcs_shutdown()


def stream():
    return _ffi.lib().pcs_stream()


def synchronize():
    _ffi.check(_ffi.lib().pcs_synchronize())

"""ndzip_b200 — B200-native implementation of ndzip's data-parallel hot path.

Host-side mirror of the reference's CUDA-facing interface (include/ndzip/cuda.hh, offload.hh,
ndzip.hh of celerity/ndzip) on top of the C ABI in include/ndzip_b200.h. The CUDA library is loaded
lazily on first use; importing the package (e.g. for ``ndzip_b200.synth``) needs neither a GPU nor
the built library.
"""
from .api import (  # noqa: F401
    NdzipB200Error,
    bind_host_to_device,
    compressed_length_bound,
    compressor_requirements,
    cuda_compressor,
    cuda_decompressor,
    cuda_offloader,
    device_numa_node,
    make_cuda_compressor,
    make_cuda_decompressor,
    make_cuda_offloader,
    num_hypercubes,
)

"""xm_code_b200 — B200-native (sm_100a) implementation of the XM Burer-Monteiro trust-region path.

Host-side Python mirror of the reference surface for this path only:

  * ``capi``      ctypes binding of the C-ABI (include/xm_b200.h, libxm_b200.so) — what tests and bench.py call
  * ``solver``    ``solve`` / ``solve_rank3`` / ``solve_rebuttle`` (XM/src/XM_main.cu:35-401): thin binding of the C-ABI's ``xm_solve``
                  (the rank staircase incl. the certificate, the same code the compiled module ``XM`` runs; block-CSR / communicator capable)
  * ``dist``      torch.distributed plumbing for one solve partitioned by camera over the GPUs of a node
  * ``recover``   ``recover_XM`` with the reference's signature (utils/recoversolution.py:4-86) on the GPU
  * ``creatematrix``  ``create_matrix`` with the reference's signature (utils/creatematrix.py:52) — Q / Abar assembled on the GPU (``xm_create_matrix``)
  * ``xm2``       the XM^2 outer loop of the pipeline scripts (3_test_colmap_glomap.py:280-350): residuals, outlier cut, graph clean-up
  * ``binio``     the ``.bin`` wire format (utils/io.py:17-58)
  * ``problems``  synthetic Q generators for the BASELINE configs (no reference code involved)

The compiled pybind11 module ``XM`` (XM/build/, built by XM/CMakeLists.txt or __graft_entry__.build()) exposes the
same three functions from C++ for the reference's demo scripts.  Nothing in this package imports ``oracle/``.
"""
from . import binio  # noqa: F401

__all__ = ["binio", "capi", "solver", "dist", "recover", "creatematrix", "xm2", "problems"]

"""cuda_gcn_b200 — B200-native full-batch GCN training path (drop-in for hengdashi/cuda_gcn's Module API).

The product is C++/CUDA: `csrc/` (sm_100a kernels behind include/gcnk.h -> libgcnk.so) and `host/`
(the C++ Module/Variable/Adam/Parser/GCN layer + `gcn-cuda` CLI -> libgcnhost.so).  The Python in this
package is only a ctypes face over those two C ABIs for the tests and bench.py.
"""

// test_host_tables.cu — prints the 1-D basis tables the CUDA kernels receive (csrc/host_tables.h: host_fill_tables) for
// every QkDG basis and degree, as JSON lines.  Host-only program (no kernel is launched), compiled with nvcc because the
// header shares definitions with device code; compared with an independent numpy evaluation by tests/test_cpp_partition.py.
#include <cstdio>
#include <cstring>
#include <vector>

#include "../../dune-pdelab_b200/csrc/common.cuh"
#include "../../dune-pdelab_b200/csrc/host_tables.h"

int main() {
  for (int basis = 0; basis <= 2; basis++)
    for (int k = 1; k <= pdb::MAX_K; k++) {
      pdb::DevParams P;
      std::memset(&P, 0, sizeof(P));
      P.dim = 2;
      P.k = k;
      P.n1 = k + 1;
      P.m = k + 1;
      P.basis = basis;
      pdb::Kron1D K;
      std::vector<double> xq, wq;
      pdb::host_fill_tables(P, K, xq, wq);
      std::printf("{\"basis\": %d, \"k\": %d, \"xq\": [", basis, k);
      for (int q = 0; q < P.m; q++) std::printf("%s%.17g", q ? ", " : "", xq[q]);
      std::printf("], \"P\": [");
      for (int i = 0; i < (P.m + 2) * P.n1; i++) std::printf("%s%.17g", i ? ", " : "", P.P[i]);
      std::printf("], \"DP\": [");
      for (int i = 0; i < (P.m + 2) * P.n1; i++) std::printf("%s%.17g", i ? ", " : "", P.DP[i]);
      std::printf("]}\n");
    }
  return 0;
}

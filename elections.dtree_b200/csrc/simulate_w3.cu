// simulate_w3.cu — kernels for ballots packed in 3 64-bit word(s).
#include "simulate_impl.cuh"
namespace dtb {
DTREE_INSTANTIATE(3)
}

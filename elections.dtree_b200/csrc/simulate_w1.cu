// simulate_w1.cu — kernels for ballots packed in 1 64-bit word(s).
#include "simulate_impl.cuh"
namespace dtb {
DTREE_INSTANTIATE(1)
}

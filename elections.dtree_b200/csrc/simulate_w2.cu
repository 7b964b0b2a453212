// simulate_w2.cu — kernels for ballots packed in 2 64-bit word(s).
#include "simulate_impl.cuh"
namespace dtb {
DTREE_INSTANTIATE(2)
}

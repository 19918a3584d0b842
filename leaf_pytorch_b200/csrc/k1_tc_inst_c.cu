// Instantiations of the tcgen05 Gabor kernel (k1_tc_kernel.cuh): forward, channel groups of 96 / 112 / 128
#include "k1_tc_kernel.cuh"
namespace leafk {
template cudaError_t launch_cg<96>(int, const Geom&, const float*, const uint8_t*, const float*, float*, int, int, int, cudaStream_t, const TcReady&, const TcMap&);
template cudaError_t launch_cg<112>(int, const Geom&, const float*, const uint8_t*, const float*, float*, int, int, int, cudaStream_t, const TcReady&, const TcMap&);
template cudaError_t launch_cg<128>(int, const Geom&, const float*, const uint8_t*, const float*, float*, int, int, int, cudaStream_t, const TcReady&, const TcMap&);
}  // namespace leafk

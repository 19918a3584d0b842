// Instantiations of the tcgen05 Gabor kernel (k1_tc_kernel.cuh): forward, channel groups of 16 / 32 / 48
#include "k1_tc_kernel.cuh"
namespace leafk {
template cudaError_t launch_cg<16>(int, const Geom&, const float*, const uint8_t*, const float*, float*, int, int, int, cudaStream_t, const TcReady&, const TcMap&);
template cudaError_t launch_cg<32>(int, const Geom&, const float*, const uint8_t*, const float*, float*, int, int, int, cudaStream_t, const TcReady&, const TcMap&);
template cudaError_t launch_cg<48>(int, const Geom&, const float*, const uint8_t*, const float*, float*, int, int, int, cudaStream_t, const TcReady&, const TcMap&);
}  // namespace leafk

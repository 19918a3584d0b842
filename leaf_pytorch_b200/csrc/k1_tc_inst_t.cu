// Instantiations of the tcgen05 Gabor kernel (k1_tc_kernel.cuh): training forward
#include "k1_tc_kernel.cuh"
namespace leafk {
template cudaError_t launch_inst<96, 3, 1, 26>(const Geom&, const float*, const uint8_t*, const float*, float*, int, int, int, cudaStream_t, const TcTrainArgs&, const TcReady&, const TcMap&);
template cudaError_t launch_inst<96, 3, 1, 0>(const Geom&, const float*, const uint8_t*, const float*, float*, int, int, int, cudaStream_t, const TcTrainArgs&, const TcReady&, const TcMap&);
template cudaError_t launch_inst<48, 5, 1, 0>(const Geom&, const float*, const uint8_t*, const float*, float*, int, int, int, cudaStream_t, const TcTrainArgs&, const TcReady&, const TcMap&);
}  // namespace leafk

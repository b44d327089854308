// sc_d3q19.cu -- explicit instantiations of stream_collide / update_fields / initialize for D3Q19
// (one translation unit per velocity set so the library builds in parallel).
// MHD on D2Q9 is not instantiated: the reference's MHD code divides by DEF_NZ/2^depth = 0 there (sim.cl:431).
#include "stream_collide_v4.cuh"
namespace ion {
ION_DEFINE_VS_LAUNCHERS(ION_D3Q19, 1)
}

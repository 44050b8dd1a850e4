// tcgen05 walkers (CIC_PREC_TC) - placeholder until the tensor-core kernels land.
#include "plan.cuh"
namespace cic {
int build_plan_tc(cic_plan*, const cic_tensor*, int, const std::string&) { set_error("CIC_PREC_TC is not available in this build"); return CIC_ERR_INVALID; }
int adaptive_forward_tc(cic_plan*, Ctx&, const cic_adaptive_io*, int, int, int) { set_error("CIC_PREC_TC is not available in this build"); return CIC_ERR_INVALID; }
int autoencoder_forward_tc(cic_plan*, Ctx&, const float*, float*, uint8_t*, int, int, int) { set_error("CIC_PREC_TC is not available in this build"); return CIC_ERR_INVALID; }
int encoder_forward_tc(cic_plan*, Ctx&, const float*, float*, float*, float*, float*, int) { set_error("CIC_PREC_TC is not available in this build"); return CIC_ERR_INVALID; }
int generator_forward_tc(cic_plan*, Ctx&, const float*, const float*, const float*, const float*, float*, int) { set_error("CIC_PREC_TC is not available in this build"); return CIC_ERR_INVALID; }
}

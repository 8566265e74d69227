// bf16 tensor-core decoder (tcgen05) - placeholder until the kernel lands.
#pragma once
#include <map>
#include <string>
#include <vector>

#include "../../include/gstk.h"
#include "common.cuh"
#include "decoder_fp32.cuh"

namespace gstk {

struct Bf16State {
  int dummy = 0;
};

inline bool bf16_config_supported(const GstkConfig&, std::string& why) {
  why = "not built yet";
  return false;
}
inline int bf16_prepare(Bf16State&, const GstkConfig&, const std::map<std::string, std::vector<float>>&, std::string& err) {
  err = "bf16 path not built";
  return GSTK_ENOTIMPL;
}
inline int bf16_decode(Bf16State&, const GstkConfig&, DecParams&, int, cudaStream_t, cudaEvent_t, cudaEvent_t, int64_t&,
                       std::string& err) {
  err = "bf16 path not built";
  return GSTK_ENOTIMPL;
}
inline void bf16_release(Bf16State&) {}

}  // namespace gstk

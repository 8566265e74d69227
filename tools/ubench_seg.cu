// Runs the decoder's real tc_segment() in isolation (diagnostics).
#include <cstdio>
#include <vector>
#include "../gst_tacotron_b200/csrc/decoder_bf16.cuh"
using namespace gstk;
namespace gstk {
template <int V>
__device__ __noinline__ uint32_t tc_segment_v(TcPipe pp, const __nv_bfloat16* act, int nkb, int wb0_base, uint32_t d0_col,
                                           bool acc0_first, int wb1_base, uint32_t d1_col, bool acc1_first,
                                           uint64_t* commit_after_wb0) {
  const int n = pp.MT * nkb;
  const uint32_t idesc = make_idesc_bf16(128, 32);
  int issued = 0, done = 0;
  const long long t0 = clock64();
  long long t_copy = 0, t_wait = 0, t_mma = 0, tl = t0;
  unsigned int spins = 0;
  while (done < n) {
    if (!(V & 4)) { while (issued < n && issued < done + TC_NSTAGE && tc_try_issue_copy(pp, act, nkb, wb0_base, wb1_base, issued)) ++issued; }
    if (!(V & 1)) { const long long now = clock64(); t_copy += now - tl; tl = now; }
    const uint32_t g = pp.it + done;
    const uint32_t s = g % TC_NSTAGE;
    if (!(V & 4) && !mbar_try_wait(&pp.full[s], (g / TC_NSTAGE) & 1u)) {
      if ((++spins & 0xFFFFu) == 0u && clock64() - t0 > 4000000000LL) __trap();
      if (!(V & 1)) { const long long now = clock64(); t_wait += now - tl; tl = now; }
      continue;
    }
    if (!(V & 1)) { const long long now = clock64(); t_wait += now - tl; tl = now; }
    if (!(V & 8)) tc_fence_after();
    const int mt = done % pp.MT, kb = (done / pp.MT + pp.rot) % nkb;
    const bool first_kb = (done / pp.MT) == 0;
    uint8_t* st = pp.stages + (size_t)s * TC_STAGE_BYTES;
    const uint64_t ad = make_desc_sw128(smem_u32(st));
    const int slot0 = tc_res_slot(wb0_base + kb);
    const uint64_t bd0 = make_desc_sw128(slot0 >= 0 ? smem_u32(pp.wres + (size_t)slot0 * TC_B_BYTES) : smem_u32(st + TC_A_BYTES));
    const int slot1 = wb1_base >= 0 ? tc_res_slot(wb1_base + kb) : 0;
    const uint64_t bd1 = make_desc_sw128(slot1 >= 0 ? smem_u32(pp.wres + (size_t)slot1 * TC_B_BYTES) : smem_u32(st + TC_A_BYTES));
    // k16 sub-step k accumulates into chain k: dependent MMAs on one accumulator are >= 4 issues apart
    const uint32_t dc0 = pp.tmem + d0_col + (uint32_t)mt * (TC_NCH * 32u);
    const uint32_t dc1 = pp.tmem + d1_col + (uint32_t)mt * (TC_NCH * 32u);
    const uint32_t accf0 = (!first_kb || !acc0_first) ? 1u : 0u;
    const uint32_t accf1 = (!first_kb || !acc1_first) ? 1u : 0u;
    if (!(V & 2) && elect_one()) {
#pragma unroll
      for (int k = 0; k < 4; ++k) umma_bf16_ss(dc0 + 32u * k, ad + 2 * k, bd0 + 2 * k, idesc, accf0);
      if (commit_after_wb0 && done == n - 1) umma_commit(commit_after_wb0);
      if (wb1_base >= 0) {
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16_ss(dc1 + 32u * k, ad + 2 * k, bd1 + 2 * k, idesc, accf1);
      }
      if (V & 32) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&pp.empty[s])) : "memory"); else umma_commit(&pp.empty[s]);
    }
    if ((V & 2) && elect_one()) { if (V & 32) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&pp.empty[s])) : "memory"); else umma_commit(&pp.empty[s]); }
    __syncwarp();
    ++done;
    if (!(V & 1)) { const long long now = clock64(); t_mma += now - tl; tl = now; }
  }
  if (pp.prof && (threadIdx.x & 31) == 0) {
    const int seg = commit_after_wb0 ? (wb1_base >= 0 ? 11 : 10) : 12;
    pp.prof[seg] += (unsigned long long)(clock64() - t0);
    pp.prof[13] += (unsigned long long)t_wait;
    pp.prof[14] += (unsigned long long)t_mma;
    pp.prof[15] += (unsigned long long)t_copy;
  }
  return pp.it + (uint32_t)n;
}

}

template <int V>
__global__ void __launch_bounds__(TC_THREADS, 1) k(const __nv_bfloat16* act, const __nv_bfloat16* wimg, long long* out, int B, int mode) {
  extern __shared__ __align__(1024) uint8_t sm_raw[];
  uint8_t* sm = (uint8_t*)(((uintptr_t)sm_raw + 1023) & ~(uintptr_t)1023);
  __shared__ __align__(8) uint64_t bars[2 * TC_NSTAGE + 3];
  __shared__ uint32_t tmem_s;
  const int tid = threadIdx.x;
  const int wid = __shfl_sync(0xffffffffu, tid >> 5, 0);
  if (tid == 0) { for (int i = 0; i < 2 * TC_NSTAGE + 3; ++i) mbar_init(&bars[i], 1); mbar_fence_init(); }
  if (wid == 0) tmem_alloc(&tmem_s, 512);
  tc_fence_before(); __syncthreads(); tc_fence_after();
  TcPipe pp;
  pp.full = bars; pp.empty = bars + TC_NSTAGE; pp.wres = sm; pp.stages = sm + (size_t)TC_RES_WB * TC_B_BYTES;
  pp.wimg_cta = (const uint8_t*)wimg; pp.it = 0; pp.tmem = tmem_s; pp.MT = (B + 127) / 128; pp.B = B; pp.rot = blockIdx.x; pp.prof = nullptr;
  uint64_t* dfull = bars + 2 * TC_NSTAGE;
  if (wid == TC_PA_WARPS) {
    long long t0 = clock64();
    for (int rep = 0; rep < 4; ++rep) {
      pp.it = tc_segment_v<V>(pp, act, TC_NKB_H, TC_WB_U2, TC_D2, true, -1, 0u, false, nullptr);          // like seg P
      pp.it = tc_segment_v<V>(pp, act, TC_NKB_X, TC_WB_W1X, TC_D1, false, -1, 0u, false, nullptr);        // like seg B (no commit)
      pp.it = tc_segment_v<V>(pp, act, TC_NKB_H, TC_WB_W2, TC_D2, false, TC_WB_U1, TC_D1, true, nullptr); // like seg C
    }
    if (elect_one()) umma_commit(dfull);
    __syncwarp();
    mbar_wait(dfull, 0);
    long long t1 = clock64();
    if ((tid & 31) == 0) out[blockIdx.x] = t1 - t0;
  } else if (mode == 1 && wid < 8) {
    mbar_wait_backoff(dfull, 0);      // like the epilogue warps
  } else if (mode == 2 && wid < 8) {
    mbar_wait(dfull, 0);              // tight spin
  }
  __syncthreads();
  if (wid == 0) tmem_dealloc(tmem_s, 512);
}


template <int V> void run(const char* name, const __nv_bfloat16* act, const __nv_bfloat16* wimg, long long* out, size_t smem) {
  cudaFuncSetAttribute(k<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  for (int B : {1, 256}) {
    for (int rep = 0; rep < 2; ++rep) k<V><<<1, TC_THREADS, smem>>>(act, wimg, out, B, 0);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return; }
    long long h; cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
    const int MT = (B + 127) / 128, units = 4 * MT * (16 + 6 + 16);
    printf("%-40s B=%3d: %6.0f ticks per unit\n", name, B, (double)h / units);
  }
}
int main() {
  __nv_bfloat16 *act, *wimg; cudaMalloc(&act, 16 * 2 * 16384); cudaMemset(act, 0, 16 * 2 * 16384);
  cudaMalloc(&wimg, TC_NWB * TC_B_BYTES); cudaMemset(wimg, 0, TC_NWB * TC_B_BYTES);
  long long* out; cudaMalloc(&out, 148 * 8);
  const size_t smem = 1024 + (size_t)TC_RES_WB * TC_B_BYTES + (size_t)TC_NSTAGE * TC_STAGE_BYTES;
  run<0>("baseline", act, wimg, out, smem);
  run<1>("no clock64", act, wimg, out, smem);
  run<1 | 2>("no clock64, no MMA", act, wimg, out, smem);
  run<1 | 4>("no clock64, no copies", act, wimg, out, smem);
  run<1 | 8>("no clock64, no tc_fence_after", act, wimg, out, smem);
  run<1 | 2 | 4>("no clock64, no MMA, no copies", act, wimg, out, smem);
  run<1 | 2 | 32>("copies only, plain arrive release", act, wimg, out, smem);
  run<1 | 4 | 32>("MMA only, plain arrive release", act, wimg, out, smem);
  run<1 | 32>("both, plain arrive release (unsafe)", act, wimg, out, smem);
  return 0;
}

/* Bring-up diagnostics of the tcgen05 act kernel (tests/tools/tc_bringup.py, tests/test_gpu_policy_parity.py).  Exported by
 * libirrl_b200.so but NOT part of the drop-in boundary (include/irrl_b200.h): nothing in the reference binds to these. */
#ifndef IRRL_DIAG_H
#define IRRL_DIAG_H
#ifdef __cplusplus
extern "C" {
#endif
/* Probe of the tcgen05 path: d[128,n] = a[128,k] b[n,k]^T (host pointers) with the 3xTF32 split of the act kernel.
 * variant bit 0 = single tf32 pass (accuracy control). */
/* Diagnostic: SM-clock timestamps of CTA (0,0) of the last tcgen05 act launch (32 slots, policy_tc_kernels.cu TC_MARK);
 * enable != 0 turns recording on for the following launches.  out32 may be NULL. */
int irrl_tc_timeline(int enable, long long* out32);
/* Diagnostic: SM cycles for `reps` back-to-back M128 x n x K8 tf32 tcgen05.mma with shared-memory operands described by
 * (layout_type, lbo, sbo) and a per-instruction start-address advance kadv; cycles2 = {issue done, all complete}. */
int irrl_tc_mma_rate(int n, int reps, unsigned layout_type, unsigned lbo, unsigned sbo, unsigned kadv, long long* cycles2);
int irrl_tc_gemm_probe(const float* a, const float* b, float* d, int k, int n, int variant);
#ifdef __cplusplus
}
#endif
#endif /* IRRL_DIAG_H */

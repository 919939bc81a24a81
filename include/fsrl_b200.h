/* fsrl_b200 -- C-ABI of the B200-native FSRL hot path.
 *
 * The reference (liuzuxin/FSRL) is pure Python and has NO FFI layer of its own
 * (SURVEY.md F1, 8b): its boundary is the Python class API (fsrl.policy / fsrl.data).  This
 * header is the boundary the new engine adds underneath that API: every entry point names
 * the reference function it replaces (file:line under /root/reference).  INTEGRATION.md
 * shows the ctypes binding a maintainer of the reference would add at each call site.
 *
 * Conventions (all entry points):
 *   - plain pointers + sizes, no torch types; every pointer is DEVICE memory unless the
 *     parameter is documented "host";
 *   - the caller owns every buffer (functions never allocate or free) and passes scratch
 *     space explicitly; `*_workspace_bytes` reports the size;
 *   - asynchronous on `stream` (a cudaStream_t passed as void*); no device sync inside;
 *   - returns 0 on success, <0 on error (FSRL_EINVAL -1, FSRL_ECUDA -2,
 *     FSRL_EWORKSPACE -3); fsrl_last_error() returns the thread-local message.  The
 *     reference signals errors with Python assert/exceptions; the Python host layer
 *     (fsrl_b200/_lib.py) re-raises these codes as the same exception types/messages;
 *   - fp32 storage; the GAE / n-step scans accumulate in fp64 like the reference.
 *   - one host thread per GPU/rank; entry points are not re-entrant on the same workspace.
 */
#ifndef FSRL_B200_H
#define FSRL_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- plumbing ------------------------------------------------------------------- */
const char* fsrl_last_error(void);
int fsrl_abi_version(void);
int fsrl_sm_count(void);

/* ---- a6/a7: dual GAE(lambda) -------------------------------------------------------
 * Replaces fsrl/policy/base_policy.py:524-540 (gae_return, numba) together with the
 * value_mask / end_flag / ret=adv+v / f32 cast of compute_gae_returns (:409-411,:429,
 * :438-446) for the reward and cost critics in ONE pass.
 *   v, vnext : [C][ld] f32   V_i(obs), V_i(obs_next)   (critic i at offset i*ld)
 *   rew,cost : [N] f32       metrics of critic 0 / 1 (cost may be NULL when C == 1)
 *   end_flag : [N] u8        terminated | truncated | unfinished-tail  (:410-411)
 *   terminated: [N] u8 or NULL; when given, vnext is masked by ~terminated (:375,:429)
 *   adv, ret : [C][ld] f32   outputs (batch.advs / batch.rets columns)
 * Flat order is the reference's batch order: env-major, chronological inside an env. */
size_t fsrl_gae_dual_workspace_bytes(int64_t N);
int fsrl_gae_dual(const float* v, const float* vnext, const float* rew, const float* cost,
                  const uint8_t* end_flag, const uint8_t* terminated, double gamma,
                  double gae_lambda, float* adv, float* ret, int64_t N, int64_t ld, int C,
                  void* workspace, size_t workspace_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* FSRL_B200_H */

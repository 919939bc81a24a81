/* Oracle (TEST INFRASTRUCTURE ONLY): plain-C restatement of the reference's two numba
 * kernels, used because a python loop over 614 400 transitions is too slow for the
 * full-size parity checks and for the cpu_baseline leg of bench.py.
 *
 *   oracle_gae_return   <- /root/reference/fsrl/policy/base_policy.py:524-540
 *   oracle_nstep_return <- /root/reference/fsrl/policy/base_policy.py:543-567
 *
 * Compile with -O2 -ffp-contract=off so that no FMA contraction changes the f64
 * rounding relative to numba's (LLVM, no fast-math) evaluation order.
 */
#include <stdint.h>
#include <stddef.h>

/* value/value_next are f32 (the critics' dtype), rew is f64, exactly the dtypes the
 * reference warms numba up with (base_policy.py:519-520). */
void oracle_gae_return(const float *value, const float *value_next, const double *rew,
                       const uint8_t *end_flag, double gamma, double gae_lambda,
                       int64_t n, double *returns)
{
    const double gl = gamma * gae_lambda;
    double gae = 0.0;
    for (int64_t i = n - 1; i >= 0; --i) {
        /* delta = rew + value_next * gamma - value          (:534) */
        double delta = (rew[i] + (double)value_next[i] * gamma) - (double)value[i];
        /* discount = (1.0 - end_flag) * (gamma * gae_lambda) (:535) */
        double discount = (1.0 - (double)(end_flag[i] != 0)) * gl;
        gae = delta + discount * gae;                       /* :538 */
        returns[i] = gae;
    }
}

/* metric f64 [buf], end_flag u8 [buf], target_q f64 [bsz*k] in/out, indices i64 [n_step*bsz] */
void oracle_nstep_return(const double *metric, const uint8_t *end_flag, double *target_q,
                         const int64_t *indices, double gamma, int64_t n_step,
                         int64_t bsz, int64_t k)
{
    for (int64_t b = 0; b < bsz; ++b) {
        double ret = 0.0;
        int64_t g = n_step;
        for (int64_t n = n_step - 1; n >= 0; --n) {
            int64_t now = indices[n * bsz + b];
            if (end_flag[now]) { g = n + 1; ret = 0.0; }
            ret = metric[now] + gamma * ret;
        }
        double gp = 1.0;
        for (int64_t i = 0; i < g; ++i) gp = gp * gamma;     /* gamma_buffer[g] (:552-554) */
        for (int64_t j = 0; j < k; ++j)
            target_q[b * k + j] = target_q[b * k + j] * gp + ret;
    }
}

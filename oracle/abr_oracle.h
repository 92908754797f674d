/* CPU oracle for the ABR chunk-step / MPC hot path — TEST INFRASTRUCTURE.
 *
 * Scalar, single-threaded C restatement of SPEC.md (which cites the reference
 * file:line of every rule).  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this library; the product
 * (abrsimulator_b200) never links or calls it.
 *
 * Parity status: MPC mode 0 (Profile R) is pinned against the unmodified
 * reference (tests/golden/mpc_ref_golden.json, mpc_ref_bulk.json: 10 409 decisions).
 * The dynamics of the chunk step (download against the square-wave trace, buffer
 * drain, rebuffering, pause gate, start-up latch, playback speed, latency) are
 * pinned by the reference's own tick loop, mechanically repaired and executed
 * (oracle/make_ref_simulator.py -> tests/golden/sim_ref_tick_golden.json, within
 * the loop's discretisation and converging with its tick).  The north-star
 * constants layered on top (RTT, payload factor, sleep quantum, per-step reward),
 * MPC mode 1, the expsmoothing predictor and the start-up phase are "parity
 * unpinned": defined by SPEC.md, no executable reference exists for them.
 *
 * Build: gcc -O2 -ffp-contract=off -fno-fast-math -shared -fPIC (oracle/Makefile).
 */
#ifndef ABR_ORACLE_H
#define ABR_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct OrcParams {          /* same field order as AbrParams in include/abr_b200.h */
    double chunk_length, max_buffer, rtt, payload, sleep_quantum;
    double rebuf_penalty, smooth_penalty, utility_scale;
    double bba_reservoir, bba_cushion;
    double start_up_length, startup_penalty, latency_penalty, latency_tick;      /* live mode, SPEC §7 */
    int32_t utility_mode, default_quality, auto_reset, hist_k;
    int32_t track_history, reserved0, live, smooth_prev_ladder;
} OrcParams;

enum { ORC_POLICY_FIXED = 0, ORC_POLICY_RANDOM = 1, ORC_POLICY_BBA = 2 };
enum { ORC_NUM_STATS = 11, ORC_NUM_ACC = 11 };  /* reward, rebuf, u, smooth, sleep, delay, steps, episodes, startup, latency integral, content played */

typedef struct OrcEnv OrcEnv;

OrcEnv* orc_env_create(const double* trace_bw, const int32_t* trace_len, const double* trace_interval,
                       int n_traces, int T_max, const double* sizes, const double* bitrates, int V, int A,
                       const OrcParams* p, int N);
void orc_env_destroy(OrcEnv* e);
void orc_env_reset(OrcEnv* e, const int32_t* trace_id, const double* start_offset);
/* one chunk step for all N sessions (SPEC §3); any output pointer may be NULL */
void orc_env_step(OrcEnv* e, const int32_t* action, double* delay, double* sleep, double* buffer,
                  double* rebuf, double* reward, double* next_sizes, uint8_t* eov, double* throughput);
/* live mode (SPEC §7): speed is the [V][N] playback-speed table (speed[k][s]: session s plays content chunk k at that
 * speed; NULL = 1) and latency the extra output; also valid with live = 0 */
void orc_env_step_live(OrcEnv* e, const int32_t* action, const double* speed, double* delay, double* sleep,
                       double* buffer, double* rebuf, double* reward, double* latency, double* next_sizes,
                       uint8_t* eov, double* throughput, double* acc /* [ORC_NUM_ACC][N] accumulated into, nullable */);
/* fused episode (SPEC §3+§4): trajectories are [steps][N]; acc is [ORC_NUM_ACC][N] */
void orc_env_rollout_live(OrcEnv* e, int policy, uint64_t seed, int64_t session_base, int steps,
                          const int32_t* actions_in, const double* speed /*[V][N] or NULL*/, double* delay,
                          double* sleep, double* buffer, double* rebuf, double* reward, double* latency, uint8_t* eov,
                          int32_t* actions_out, double* acc);
void orc_env_rollout(OrcEnv* e, int policy, uint64_t seed, int64_t session_base, int steps,
                     const int32_t* actions_in, double* delay, double* sleep, double* buffer, double* rebuf,
                     double* reward, uint8_t* eov, int32_t* actions_out, double* acc);
/* MPC over the env's own state and history ring (SPEC §5); mode 0 = Profile R, 1 = robust */
void orc_env_mpc_decide(OrcEnv* e, int H, int mode, int32_t* action, double* best_j, int32_t* best_seq);
/* statistics vector from an acc table (SPEC §6): plain ascending-session sums */
void orc_stats_from_acc(const double* acc, int N, double* out);
/* state access for tests: field 0 seg,1 chunk,2 last_q,3 trace_id,4 hist_len (int32) ; 10 phase,11 buffer (double);
 * live mode: 16 t_now, 17 play_time, 19 play_len (double), 7 started (uint8), 8 play_id (int32) */
const void* orc_env_field(OrcEnv* e, int field);
int orc_env_error_count(OrcEnv* e);

/* standalone batched MPC decision (SPEC §5).  History is a per-session ring [N][K]
 * (slot (hist_len-1) mod K newest; hist_len <= K means plain oldest-first rows).
 * last_pred/err_ring/err_len ([N], [N][K], [N]) may be NULL (no robust error state).
 * Outputs: action[N] (-1 on input error), best_j[N], best_seq[N][H], preds[N][H] (mode 0) — nullable. */
void orc_mpc_decide(const double* sizes, const double* util, int V, int A, const OrcParams* p, int N,
                    const int32_t* chunk_idx, const int32_t* prev_q, const double* buffer,
                    const double* bw_hist, const int32_t* hist_len, int K,
                    double* last_pred, double* err_ring, int32_t* err_len,
                    int H, int mode, int32_t* action, double* best_j, int32_t* best_seq, double* preds,
                    int32_t* n_errors);

/* the same with the two off-default features of SPEC 5.3 / 5.4: ses != 0 selects the "expsmoothing" predictor of
 * mpc.py:72-79 (mode 0 only); n_ts > 1 adds the start-up delay T_s = jt * ts_step as a decision variable for the
 * sessions with startup[s] != 0 (NULL = all), returned in startup_delay[N] */
void orc_mpc_decide_ex(const double* sizes, const double* util, int V, int A, const OrcParams* p, int N,
                       const int32_t* chunk_idx, const int32_t* prev_q, const double* buffer,
                       const double* bw_hist, const int32_t* hist_len, int K,
                       double* last_pred, double* err_ring, int32_t* err_len,
                       int H, int mode, int ses, const uint8_t* startup, int n_ts, double ts_step,
                       int32_t* action, double* startup_delay, double* best_j, int32_t* best_seq, double* preds,
                       int32_t* n_errors);

/* helpers exposed for unit tests */
void orc_philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t out[4]);
void orc_utility_table(const double* bitrates, int V, int A, int mode, double scale, double* util);

#ifdef __cplusplus
}
#endif
#endif

/* CPU oracle — TEST INFRASTRUCTURE (see abr_oracle.h for the rules of use).
 *
 * Every function restates one section of SPEC.md, which in turn cites the
 * reference lines (Simulator.py / mpc.py).  Deliberately scalar and naive: each
 * MPC sequence is rolled out independently from scratch exactly like
 * mpc.py:120-162 does, with no prefix sharing.
 */
#include "abr_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define MAXH 10

struct OrcEnv {
    int n_traces, T_max, V, A, N, K;
    OrcParams p;
    double *trace_bw, *trace_interval, *sizes, *bitrates, *util;
    double *cum /* [n_traces][T_max + 1], SPEC 3.1 */;
    int32_t* trace_len;
    /* session state, SoA (SPEC §1) */
    int32_t *seg, *chunk, *last_q, *trace_id, *hist_len, *err_len;
    double *phi /* fraction of segment seg consumed */, *pos /* the same position in data coordinates, SPEC 3.1 */, *buffer, *bw_hist /* [N][K] ring */, *last_pred, *err_ring /* [N][K] */;
    double *t_now, *play_time, *play_len;   /* live mode, SPEC §7 */
    int32_t* play_id;
    uint8_t *done, *started;
    int errors, bad_speed;
    uint32_t step_base;   /* fused-episode steps since the last reset (SPEC §4: step_index of the random policy) */
};

static inline double max0(double x) { return x > 0.0 ? x : 0.0; } /* Python max(0, x), mpc.py:107 */

void orc_utility_table(const double* bitrates, int V, int A, int mode, double scale, double* util) {
    for (int v = 0; v < V; ++v)
        for (int a = 0; a < A; ++a) {
            double b = bitrates[v * A + a];
            util[v * A + a] = (mode == 1) ? log(b / bitrates[v * A + A - 1]) /* mpc.py:99-102 */
                                          : b * scale;                      /* mpc.py:95-97 */
        }
}

/* ---------------- Philox4x32-10 (SPEC §4) ---------------- */
void orc_philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t out[4]) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)M0 * c0, p1 = (uint64_t)M1 * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += W0; k1 += W1;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* ---------------- env ---------------- */
static void* dup_mem(const void* src, size_t bytes) {
    void* d = malloc(bytes ? bytes : 1);
    if (src) memcpy(d, src, bytes);
    return d;
}

OrcEnv* orc_env_create(const double* trace_bw, const int32_t* trace_len, const double* trace_interval,
                       int n_traces, int T_max, const double* sizes, const double* bitrates, int V, int A,
                       const OrcParams* p, int N) {
    OrcEnv* e = (OrcEnv*)calloc(1, sizeof(OrcEnv));
    e->n_traces = n_traces; e->T_max = T_max; e->V = V; e->A = A; e->N = N; e->p = *p;
    e->K = p->hist_k > 0 ? p->hist_k : 1;
    e->trace_bw = (double*)dup_mem(trace_bw, sizeof(double) * (size_t)n_traces * T_max);
    e->trace_len = (int32_t*)dup_mem(trace_len, sizeof(int32_t) * n_traces);
    e->trace_interval = (double*)dup_mem(trace_interval, sizeof(double) * n_traces);
    e->sizes = (double*)dup_mem(sizes, sizeof(double) * V * A);
    e->bitrates = (double*)dup_mem(bitrates, sizeof(double) * V * A);
    e->util = (double*)malloc(sizeof(double) * V * A);
    orc_utility_table(bitrates, V, A, p->utility_mode, p->utility_scale, e->util);
    /* SPEC 3.1 table: C[0] = 0 ; C[j+1] = C[j] + (bw[j]*payload)*I, left to right */
    e->cum = (double*)calloc((size_t)n_traces * (T_max + 1), 8);
    for (int t = 0; t < n_traces; ++t) {
        const double I = trace_interval[t];
        double c = 0.0;
        e->cum[(size_t)t * (T_max + 1)] = 0.0;
        for (int j = 0; j < trace_len[t]; ++j) {
            double r = trace_bw[(size_t)t * T_max + j] * p->payload;
            c = c + r * I;
            e->cum[(size_t)t * (T_max + 1) + j + 1] = c;
        }
    }
    e->seg = (int32_t*)calloc(N, 4); e->chunk = (int32_t*)calloc(N, 4); e->last_q = (int32_t*)calloc(N, 4);
    e->trace_id = (int32_t*)calloc(N, 4); e->hist_len = (int32_t*)calloc(N, 4); e->err_len = (int32_t*)calloc(N, 4);
    e->phi = (double*)calloc(N, 8); e->pos = (double*)calloc(N, 8); e->buffer = (double*)calloc(N, 8); e->last_pred = (double*)calloc(N, 8);
    e->bw_hist = (double*)calloc((size_t)N * e->K, 8); e->err_ring = (double*)calloc((size_t)N * e->K, 8);
    e->done = (uint8_t*)calloc(N, 1);
    e->started = (uint8_t*)calloc(N, 1);
    e->t_now = (double*)calloc(N, 8); e->play_time = (double*)calloc(N, 8);
    e->play_len = (double*)calloc(N, 8); e->play_id = (int32_t*)calloc(N, 4);
    return e;
}

void orc_env_destroy(OrcEnv* e) {
    if (!e) return;
    free(e->cum);
    free(e->trace_bw); free(e->trace_len); free(e->trace_interval); free(e->sizes); free(e->bitrates); free(e->util);
    free(e->seg); free(e->chunk); free(e->last_q); free(e->trace_id); free(e->hist_len); free(e->err_len);
    free(e->phi); free(e->pos); free(e->buffer); free(e->last_pred); free(e->bw_hist); free(e->err_ring); free(e->done);
    free(e->started); free(e->t_now); free(e->play_time); free(e->play_len); free(e->play_id);
    free(e);
}

const void* orc_env_field(OrcEnv* e, int f) {
    switch (f) {
        case 0: return e->seg; case 1: return e->chunk; case 2: return e->last_q; case 3: return e->trace_id;
        case 4: return e->hist_len; case 5: return e->done; case 6: return e->err_len;
        case 10: return e->phi; case 11: return e->buffer; case 12: return e->bw_hist; case 13: return e->last_pred;
        case 14: return e->err_ring; case 15: return e->util; case 16: return e->t_now; case 17: return e->play_time;
        case 7: return e->started; case 18: return e->pos; case 8: return e->play_id; case 19: return e->play_len;
    }
    return 0;
}
int orc_env_error_count(OrcEnv* e) { return e->errors; }

/* SPEC §2 */
void orc_env_reset(OrcEnv* e, const int32_t* trace_id, const double* start_offset) {
    e->step_base = 0;
    for (int s = 0; s < e->N; ++s) {
        int tr = trace_id[s];
        int T = e->trace_len[tr];
        double I = e->trace_interval[tr];
        double off = start_offset ? start_offset[s] : 0.0;
        double x = off / I;
        double n = floor(x);
        int seg = (int)fmod(n, (double)T);
        double phi = x - n;
        const double* C = e->cum + (size_t)tr * (e->T_max + 1);
        e->trace_id[s] = tr; e->seg[s] = seg; e->phi[s] = phi;
        e->pos[s] = C[seg] + (C[seg + 1] - C[seg]) * phi;
        e->buffer[s] = 0.0; e->chunk[s] = 0; e->last_q[s] = e->p.default_quality; e->done[s] = 0;
        e->hist_len[s] = 0; e->last_pred[s] = 0.0; e->err_len[s] = 0;
        e->t_now[s] = 0.0; e->play_time[s] = 0.0; e->started[s] = e->p.start_up_length <= 0.0;
        e->play_id[s] = 0; e->play_len[s] = 0.0;
    }
}

typedef struct StepOut { double delay, sleep, buffer, rebuf, reward, thr, u, smooth, latency, startup, area, played; uint8_t eov; int inert; } StepOut;

/* ---- SPEC §7 playback model: closed form of the reference's playback block, Simulator.py:174-187 ---- */
typedef struct LiveAcc { double startup, area, played, tc; } LiveAcc;   /* tc: wall clock inside the step */

/* speed of the content chunk session s is playing: speed[k][s] of the [V][N] table (Simulator.py:176-177: the speed
 * controller is asked when a chunk starts to play); a non-positive or NaN entry counts as an error and plays at 1 */
static double live_speed(OrcEnv* e, int s, const double* speed) {
    if (!speed) return 1.0;
    int k = e->play_id[s] < e->V ? e->play_id[s] : e->V - 1;
    double v = speed[(size_t)k * e->N + s];
    if (!(v > 0.0)) { e->bad_speed = 1; v = 1.0; }
    return v;
}

/* one stretch: d seconds of content in dw seconds of wall time at speed v; the latency (wall clock - content
 * played) changes at the rate 1 - v inside the stretch, area integrates it (Simulator.py:179-180) */
static void live_piece(OrcEnv* e, int s, double* buffer, LiveAcc* a, double d, double dw, double v) {
    a->area = a->area + ((a->tc - e->play_time[s]) * dw + ((1.0 - v) * dw) * (dw * 0.5));
    e->play_time[s] = e->play_time[s] + d;
    *buffer = *buffer - d;
    a->tc = a->tc + dw;
    a->played = a->played + d;
}

/* SPEC §7 play_wall(dt): playback during dt seconds of wall time; returns the stall time */
static double play_wall(OrcEnv* e, int s, const double* speed, double* buffer, LiveAcc* a, double dt) {
    const double L = e->p.chunk_length;
    if (!e->started[s]) { a->startup = a->startup + dt; a->tc = a->tc + dt; return 0.0; }
    double rem = dt;
    while (rem > 0.0 && *buffer > 0.0) {
        double v = live_speed(e, s, speed);
        double room = L - e->play_len[s];
        double can = room < *buffer ? room : *buffer;
        double need = v * rem;
        if (need < can) {
            live_piece(e, s, buffer, a, need, rem, v);
            e->play_len[s] = e->play_len[s] + need;
            rem = 0.0;
        } else {
            double dw = can / v;
            int finished = room <= *buffer;          /* the chunk ends before the buffer does */
            live_piece(e, s, buffer, a, can, dw, v);
            rem = rem - dw;
            if (finished) { e->play_id[s] += 1; e->play_len[s] = 0.0; }
            else e->play_len[s] = e->play_len[s] + can;
        }
    }
    if (rem < 0.0) rem = 0.0;
    a->tc = a->tc + rem;
    return rem;
}

/* SPEC §7 play_content(x): playback until x seconds of content have drained (x <= buffer); returns the wall time */
static double play_content(OrcEnv* e, int s, const double* speed, double* buffer, LiveAcc* a, double x) {
    const double L = e->p.chunk_length;
    double w = 0.0;
    while (x > 0.0) {
        double v = live_speed(e, s, speed);
        double room = L - e->play_len[s];
        int finished = room <= x;
        double d = finished ? room : x;
        double dw = d / v;
        live_piece(e, s, buffer, a, d, dw, v);
        w = w + dw;
        if (finished) { e->play_id[s] += 1; e->play_len[s] = 0.0; x = x - d; }
        else { e->play_len[s] = e->play_len[s] + d; x = 0.0; }
    }
    return w;
}

static void advance_trace(int* seg, double* phi, double dt, double I, int T) {   /* SPEC §3.3 */
    double x = *phi + dt / I;
    double n = floor(x);
    *phi = x - n;
    *seg = (int)((*seg + (int64_t)fmod(n, (double)T)) % T);
}

/* SPEC §3 (and §7 when live = 1) for one session; speed: [V][N] playback-speed table or NULL */
static void step_one(OrcEnv* e, int s, int q, const double* speed, StepOut* o) {
    const OrcParams* p = &e->p;
    memset(o, 0, sizeof(*o));
    if (e->done[s]) { o->eov = 1; o->buffer = e->buffer[s]; o->inert = 1; return; }
    const int tr = e->trace_id[s];
    const int T = e->trace_len[tr];
    const double I = e->trace_interval[tr];
    int chunk = e->chunk[s], seg = e->seg[s];
    double phi = e->phi[s], pos = e->pos[s], buffer = e->buffer[s];
    const double size = e->sizes[chunk * e->A + q];
    const double* C = e->cum + (size_t)tr * (e->T_max + 1);
    const int live = p->live != 0;
    double idle = 0.0, rebuf = 0.0, latency = 0.0;
    LiveAcc a = {0.0, 0.0, 0.0, e->t_now[s]};
    e->bad_speed = 0;
    if (live) {   /* 7.1 pause gate */
        double w1 = (double)(chunk + 1) * p->chunk_length - e->t_now[s];
        w1 = max0(w1);
        rebuf = play_wall(e, s, speed, &buffer, &a, w1);
        double w2 = (e->started[s] && buffer > p->max_buffer) ? play_content(e, s, speed, &buffer, &a, buffer - p->max_buffer) : 0.0;
        idle = w1 + w2;
        if (idle > 0.0) { advance_trace(&seg, &phi, idle, I, T); pos = C[seg] + (C[seg + 1] - C[seg]) * phi; }
    }
    /* 3.1 download against the cumulative capacity of the trace */
    const double P = C[T];
    double delay;
    {
        double target = pos + size;
        long long n = 0;               /* whole trace periods */
        while (target >= P) {
            target = target - P;
            n += 1;
            if (n >= (1 << 20)) { e->errors++; target = 0.0; break; }   /* safety net only: P > 0 */
        }
        /* the largest j in [0, T) with C[j] <= target; C is increasing and, without a wrap,
         * C[seg] <= pos <= target, so the scan may start at seg */
        int j = (n == 0) ? seg : 0;
        while (j + 1 < T && C[j + 1] <= target) ++j;
        double phi_new = (target - C[j]) / (C[j + 1] - C[j]);   /* fraction of segment j consumed */
        long long k = (long long)(j - seg) + n * (long long)T;   /* segment boundaries crossed */
        delay = max0(((double)k + (phi_new - phi)) * I) + p->rtt;
        seg = j;
        phi = phi_new;
        pos = target;
    }
    double thr = size / delay;
    double sleep = 0.0;
    if (live) {   /* 7.2 */
        rebuf = rebuf + play_wall(e, s, speed, &buffer, &a, delay);
        buffer = buffer + p->chunk_length;
        e->t_now[s] = (e->t_now[s] + idle) + delay;
        if (!e->started[s] && buffer >= p->start_up_length) e->started[s] = 1;
        latency = e->t_now[s] - e->play_time[s];
        sleep = idle;
    } else {
        /* 3.2 */
        rebuf = max0(delay - buffer);
        buffer = max0(buffer - delay) + p->chunk_length;
        /* 3.3 */
        if (buffer > p->max_buffer) {
            sleep = ceil((buffer - p->max_buffer) / p->sleep_quantum) * p->sleep_quantum;
            buffer = buffer - sleep;
            advance_trace(&seg, &phi, sleep, I, T);
            pos = C[seg] + (C[seg + 1] - C[seg]) * phi;
        }
    }
    /* 3.4 */
    const double u = e->util[chunk * e->A + q];
    const int lq = e->last_q[s];
    /* previous index in the current chunk's ladder (mpc.py:148-149) or, smooth_prev_ladder = 1, in the previous chunk's
     * own ladder (Simulator.calculate_qoe, Simulator.py:81-82) */
    const int prow = (p->smooth_prev_ladder && chunk > 0) ? chunk - 1 : chunk;
    double smooth = (lq >= 0) ? fabs(u - e->util[prow * e->A + lq]) : 0.0;
    double reward = (u - p->rebuf_penalty * rebuf) - p->smooth_penalty * smooth;
    if (live) reward = reward - p->latency_penalty * latency;
    /* history ring */
    if (p->track_history) {
        e->bw_hist[(size_t)s * e->K + (e->hist_len[s] % e->K)] = thr;
        e->hist_len[s] += 1;
    }
    /* 3.5 */
    chunk += 1;
    o->delay = delay; o->sleep = sleep; o->buffer = buffer; o->rebuf = rebuf; o->reward = reward;
    o->thr = thr; o->u = u; o->smooth = smooth; o->latency = latency; o->startup = a.startup; o->area = a.area; o->played = a.played;
    if (live && e->bad_speed) e->errors++;
    o->eov = (chunk >= e->V);
    e->last_q[s] = q;
    if (o->eov && p->auto_reset) {
        chunk = 0; buffer = 0.0; e->last_q[s] = p->default_quality;
        e->hist_len[s] = 0; e->last_pred[s] = 0.0; e->err_len[s] = 0;
        e->t_now[s] = 0.0; e->play_time[s] = 0.0; e->started[s] = p->start_up_length <= 0.0;
        e->play_id[s] = 0; e->play_len[s] = 0.0;
    } else if (o->eov) {
        e->done[s] = 1;
    }
    e->chunk[s] = chunk; e->seg[s] = seg; e->phi[s] = phi; e->pos[s] = pos; e->buffer[s] = buffer;
}

void orc_env_step_live(OrcEnv* e, const int32_t* action, const double* speed, double* delay, double* sleep,
                       double* buffer, double* rebuf, double* reward, double* latency, double* next_sizes,
                       uint8_t* eov, double* throughput, double* acc) {
    const size_t N = (size_t)e->N;
    for (int s = 0; s < e->N; ++s) {
        StepOut o;
        step_one(e, s, action[s], speed, &o);
        if (delay) delay[s] = o.delay;
        if (sleep) sleep[s] = o.sleep;
        if (buffer) buffer[s] = o.buffer;
        if (rebuf) rebuf[s] = o.rebuf;
        if (reward) reward[s] = o.reward;
        if (latency) latency[s] = o.latency;
        if (eov) eov[s] = o.eov;
        if (throughput) throughput[s] = o.thr;
        if (next_sizes)
            for (int a = 0; a < e->A; ++a)
                next_sizes[(size_t)s * e->A + a] = e->done[s] ? 0.0 : e->sizes[e->chunk[s] * e->A + a];
        if (acc && !o.inert) {
            acc[0 * N + s] += o.reward; acc[1 * N + s] += o.rebuf; acc[2 * N + s] += o.u; acc[3 * N + s] += o.smooth;
            acc[4 * N + s] += o.sleep; acc[5 * N + s] += o.delay; acc[6 * N + s] += 1.0;
            if (o.eov) acc[7 * N + s] += 1.0;
            acc[8 * N + s] += o.startup; acc[9 * N + s] += o.area; acc[10 * N + s] += o.played;
        }
    }
}

void orc_env_step(OrcEnv* e, const int32_t* action, double* delay, double* sleep, double* buffer,
                  double* rebuf, double* reward, double* next_sizes, uint8_t* eov, double* throughput) {
    orc_env_step_live(e, action, 0, delay, sleep, buffer, rebuf, reward, 0, next_sizes, eov, throughput, 0);
}

/* SPEC §4 */
static int policy_action(OrcEnv* e, int s, int policy, uint64_t seed, int64_t session_base, int step,
                         const int32_t* actions_in) {
    const int A = e->A;
    if (policy == ORC_POLICY_FIXED) return actions_in[(size_t)step * e->N + s];
    if (policy == ORC_POLICY_RANDOM) {
        uint64_t g = (uint64_t)(session_base + s);
        uint32_t r[4];
        /* one Philox block per eight steps: counter (session, step / 8), 16-bit slice step % 8 (low half of word 0 first) */
        const uint32_t tg = e->step_base + (uint32_t)step;   /* step index since the last reset */
        orc_philox4x32_10((uint32_t)g, (uint32_t)(g >> 32), tg >> 3, 0u, (uint32_t)seed,
                          (uint32_t)(seed >> 32), r);
        const uint32_t x16 = (r[(tg & 7) >> 1] >> (16 * (tg & 1))) & 0xffffu;
        return (int)((x16 * (uint32_t)A) >> 16);
    }
    /* BBA */
    double b = e->buffer[s];
    if (b < e->p.bba_reservoir) return 0;
    if (b >= e->p.bba_reservoir + e->p.bba_cushion) return A - 1;
    int q = (int)floor(((double)(A - 1) * (b - e->p.bba_reservoir)) / e->p.bba_cushion);
    return q > A - 1 ? A - 1 : q;
}

void orc_env_rollout_live(OrcEnv* e, int policy, uint64_t seed, int64_t session_base, int steps,
                          const int32_t* actions_in, const double* speed /*[V][N] or NULL*/, double* delay,
                          double* sleep, double* buffer, double* rebuf, double* reward, double* latency, uint8_t* eov,
                          int32_t* actions_out, double* acc) {
    const int N = e->N;
    for (int s = 0; s < N; ++s) {
        double a_rew = 0, a_reb = 0, a_u = 0, a_sm = 0, a_sl = 0, a_dl = 0, a_steps = 0, a_eps = 0, a_su = 0, a_lat = 0, a_pl = 0;
        for (int t = 0; t < steps; ++t) {
            int q = policy_action(e, s, policy, seed, session_base, t, actions_in);
            StepOut o;
            size_t ix = (size_t)t * N + s;
            step_one(e, s, q, speed, &o);
            if (delay) delay[ix] = o.delay;
            if (sleep) sleep[ix] = o.sleep;
            if (buffer) buffer[ix] = o.buffer;
            if (rebuf) rebuf[ix] = o.rebuf;
            if (reward) reward[ix] = o.reward;
            if (latency) latency[ix] = o.latency;
            if (eov) eov[ix] = o.eov;
            if (actions_out) actions_out[ix] = q;
            if (!o.inert) {
                a_rew = a_rew + o.reward; a_reb = a_reb + o.rebuf; a_u = a_u + o.u;
                a_sm = a_sm + o.smooth; a_sl = a_sl + o.sleep; a_dl = a_dl + o.delay;
                a_su = a_su + o.startup; a_lat = a_lat + o.area; a_pl = a_pl + o.played;
                a_steps += 1.0; if (o.eov) a_eps += 1.0;
            }
        }
        if (acc) {
            acc[0 * (size_t)N + s] = a_rew; acc[1 * (size_t)N + s] = a_reb; acc[2 * (size_t)N + s] = a_u;
            acc[3 * (size_t)N + s] = a_sm; acc[4 * (size_t)N + s] = a_sl; acc[5 * (size_t)N + s] = a_dl;
            acc[6 * (size_t)N + s] = a_steps; acc[7 * (size_t)N + s] = a_eps;
            acc[8 * (size_t)N + s] = a_su; acc[9 * (size_t)N + s] = a_lat; acc[10 * (size_t)N + s] = a_pl;   /* 0 outside live mode */
        }
    }
    e->step_base += (uint32_t)(steps > 0 ? steps : 0);
}

void orc_env_rollout(OrcEnv* e, int policy, uint64_t seed, int64_t session_base, int steps,
                     const int32_t* actions_in, double* delay, double* sleep, double* buffer, double* rebuf,
                     double* reward, uint8_t* eov, int32_t* actions_out, double* acc) {
    orc_env_rollout_live(e, policy, seed, session_base, steps, actions_in, 0, delay, sleep, buffer, rebuf, reward, 0,
                         eov, actions_out, acc);
}

void orc_stats_from_acc(const double* acc, int N, double* out) {
    for (int j = 0; j < ORC_NUM_STATS; ++j) {
        out[j] = 0.0;
        for (int s = 0; s < N; ++s) out[j] += acc[(size_t)j * N + s];
    }
}

/* ---------------- MPC (SPEC §5) ---------------- */
static double rollout_j(const double* U /*[H][A]*/, const double* RB, const double* DL, int A, int h, const int* R,
                        int prev_q, double buffer, int clamp, double L, double B, double vw, double rw) {
    double vq = 0.0, qv = 0.0, rt = 0.0, b = buffer;
    int ap = prev_q;
    for (int i = 0; i < h; ++i) {
        int a = R[i];
        vq = vq + U[i * A + a];
        if (ap >= 0) qv = qv + fabs(U[i * A + a] - U[i * A + ap]);
        double d = RB[i * A + a] - b;
        rt = rt + (clamp ? max0(d) : d);
        if (i != h - 1) {
            double t = max0(b - DL[i * A + a]);
            double w = max0((t + L) - B);
            b = max0((t + L) - w);
        }
        ap = a;
    }
    return -((vq - vw * qv) - rw * rt);
}

/* hist: n samples oldest first.  Returns 1 on input error (action = -1). */
static int mpc_one(const double* sizes, const double* util, int V, int A, const OrcParams* p, int k, int prev_q,
                   double buffer, const double* hist, int n, int H, int mode, double* last_pred, double* err_ring,
                   int32_t* err_len, int K, int32_t* action, double* best_j, int32_t* best_seq, double* preds_out,
                   int ses, int n_ts, double ts_step, double* ts_out) {
    double U[MAXH * 16], RB[MAXH * 16], DL[MAXH * 16];
    const double L = p->chunk_length, B = p->max_buffer;
    int h = H;
    *action = -1;
    if (ts_out) *ts_out = 0.0;
    if (best_j) *best_j = NAN;
    if (best_seq) for (int i = 0; i < H; ++i) best_seq[i] = -1;
    if (H < 1 || H > MAXH || A > 16 || k < 0) return 1;
    if (mode == 0) {
        if (n <= 0 || k + H > V || prev_q >= A) return 1; /* D14 / D13; prev_q < 0 = no previous chunk */
        double S = 0.0;
        for (int j = 0; j < n; ++j) {
            if (hist[j] == 0.0) return 1;                                  /* ZeroDivisionError, mpc.py:88 */
            S = S + 1.0 / hist[j];
        }
        double l_n = 0.0;
        if (ses) {   /* SPEC 5.4: predictor "expsmoothing", mpc.py:72-79 — flat forecast of simple exponential smoothing,
                        alpha = 0.5, initial level = least-squares fit of the one-step errors (l_t = la + lb * l_0) */
            double la = 0.0, lb = 1.0, num = 0.0, den = 0.0;
            for (int j = 0; j < n; ++j) {
                double r = hist[j] - la;
                num = num + lb * r;
                den = den + lb * lb;
                la = 0.5 * hist[j] + 0.5 * la;
                lb = 0.5 * lb;
            }
            l_n = la + lb * (num / den);
            if (!(l_n > 0.0)) return 1;
        }
        for (int i = 0; i < H; ++i) {                                      /* mpc.py:83-92 incl. D10 */
            double pi = ses ? l_n : (double)(n + i) / S;
            S = S + 1.0 / pi;
            if (preds_out) preds_out[i] = pi;
            for (int a = 0; a < A; ++a) {
                double sz = sizes[(k + i) * A + a];
                double m = sz > 0.0 ? sz : 0.0;                            /* max(0, size, L), mpc.py:151 */
                if (L > m) m = L;
                RB[i * A + a] = m / pi;
                DL[i * A + a] = sizes[k * A + a] / pi;                     /* D12: chunk k's sizes */
                U[i * A + a] = util[(k + i) * A + a];
            }
        }
    } else {
        if (prev_q >= A) return 1;
        if (n <= 0) { *action = p->default_quality > 0 ? p->default_quality : 0; return 0; }
        double S = 0.0;
        for (int j = 0; j < n; ++j) {
            if (!(hist[j] > 0.0)) return 1;
            S = S + 1.0 / hist[j];
        }
        double hm = (double)n / S;
        double max_err = 0.0;
        if (last_pred) {
            if (*last_pred > 0.0) {
                double x = hist[n - 1];
                err_ring[*err_len % K] = fabs(*last_pred - x) / x;
                *err_len += 1;
            }
            int m = *err_len < K ? *err_len : K;
            for (int j = 0; j < m; ++j) if (err_ring[j] > max_err) max_err = err_ring[j];
            *last_pred = hm;
        }
        double c = hm / (1.0 + max_err);
        if (preds_out) for (int i = 0; i < H; ++i) preds_out[i] = c;
        h = H < V - k ? H : V - k;
        if (h <= 0) { *action = 0; return 0; }
        for (int i = 0; i < h; ++i)
            for (int a = 0; a < A; ++a) {
                RB[i * A + a] = DL[i * A + a] = sizes[(k + i) * A + a] / c;
                U[i * A + a] = util[(k + i) * A + a];
            }
    }
    /* SPEC 5.3: in the start-up phase the start-up delay T_s = jt * ts_step is a second decision variable, slowest axis */
    double bj = 0.0;
    int have = 0;
    if (n_ts < 1) n_ts = 1;
    for (int jt = 0; jt < n_ts; ++jt) {
        const double ts = jt == 0 ? 0.0 : (double)jt * ts_step;
        const double b0 = jt == 0 ? buffer : buffer + ts;
        int R[MAXH] = {0};
        double bj_t = 0.0;
        int have_t = 0, act_t = -1, seq_t[MAXH];
        for (;;) {
            double j = rollout_j(U, RB, DL, A, h, R, prev_q, b0, mode != 0, L, B, p->smooth_penalty, p->rebuf_penalty);
            if (!have_t || j < bj_t) {                                         /* first minimum, C order */
                have_t = 1; bj_t = j; act_t = R[0];
                for (int i = 0; i < h; ++i) seq_t[i] = R[i];
            }
            int d = h - 1;
            while (d >= 0 && ++R[d] == A) { R[d] = 0; --d; }
            if (d < 0) break;
        }
        /* q = -J ; total = q - sw * T_s ; strict improvement keeps the smaller T_s on ties */
        const double tot = n_ts == 1 ? bj_t : -((-bj_t) - p->startup_penalty * ts);
        if (!have || tot < bj) {
            have = 1; bj = tot; *action = act_t;
            if (ts_out) *ts_out = ts;
            if (best_seq) for (int i = 0; i < h; ++i) best_seq[i] = seq_t[i];
        }
    }
    if (best_j) *best_j = bj;
    return 0;
}

static int gather_hist(const double* ring, int len, int K, double* out) {
    int n = len < K ? len : K;
    int start = len <= K ? 0 : len % K;
    for (int j = 0; j < n; ++j) out[j] = ring[(start + j) % K];
    return n;
}

void orc_mpc_decide_ex(const double* sizes, const double* util, int V, int A, const OrcParams* p, int N,
                       const int32_t* chunk_idx, const int32_t* prev_q, const double* buffer,
                       const double* bw_hist, const int32_t* hist_len, int K,
                       double* last_pred, double* err_ring, int32_t* err_len,
                       int H, int mode, int ses, const uint8_t* startup, int n_ts, double ts_step,
                       int32_t* action, double* startup_delay, double* best_j, int32_t* best_seq, double* preds,
                       int32_t* n_errors) {
    double* tmp = (double*)malloc(sizeof(double) * (K > 0 ? K : 1));
    int errs = 0;
    for (int s = 0; s < N; ++s) {
        int n = gather_hist(bw_hist + (size_t)s * K, hist_len[s], K, tmp);
        int nts = (n_ts > 1 && (!startup || startup[s])) ? n_ts : 1;
        errs += mpc_one(sizes, util, V, A, p, chunk_idx[s], prev_q[s], buffer[s], tmp, n, H, mode,
                        last_pred ? last_pred + s : 0, err_ring ? err_ring + (size_t)s * K : 0,
                        err_len ? err_len + s : 0, K, action + s, best_j ? best_j + s : 0,
                        best_seq ? best_seq + (size_t)s * H : 0, preds ? preds + (size_t)s * H : 0,
                        ses, nts, ts_step, startup_delay ? startup_delay + s : 0);
    }
    free(tmp);
    if (n_errors) *n_errors = errs;
}

void orc_mpc_decide(const double* sizes, const double* util, int V, int A, const OrcParams* p, int N,
                    const int32_t* chunk_idx, const int32_t* prev_q, const double* buffer,
                    const double* bw_hist, const int32_t* hist_len, int K,
                    double* last_pred, double* err_ring, int32_t* err_len,
                    int H, int mode, int32_t* action, double* best_j, int32_t* best_seq, double* preds,
                    int32_t* n_errors) {
    orc_mpc_decide_ex(sizes, util, V, A, p, N, chunk_idx, prev_q, buffer, bw_hist, hist_len, K, last_pred, err_ring,
                      err_len, H, mode, 0, 0, 1, 0.0, action, 0, best_j, best_seq, preds, n_errors);
}

void orc_env_mpc_decide(OrcEnv* e, int H, int mode, int32_t* action, double* best_j, int32_t* best_seq) {
    double tmp[64];
    for (int s = 0; s < e->N; ++s) {
        if (e->done[s]) { action[s] = 0; if (best_j) best_j[s] = NAN; continue; }
        int n = gather_hist(e->bw_hist + (size_t)s * e->K, e->hist_len[s], e->K, tmp);
        if (mode == 0 && n == 0) {               /* env flow: no sample yet -> default quality (SPEC §5.2 rule reused) */
            action[s] = e->p.default_quality > 0 ? e->p.default_quality : 0; if (best_j) best_j[s] = NAN; continue;
        }
        if (mode == 0 && e->chunk[s] + H > e->V) { /* env flow never raises: truncate like mode 1 */
            int h = e->V - e->chunk[s];
            e->errors += mpc_one(e->sizes, e->util, e->V, e->A, &e->p, e->chunk[s], e->last_q[s], e->buffer[s], tmp, n,
                                 h, 0, 0, 0, 0, e->K, action + s, best_j ? best_j + s : 0, 0, 0, 0, 1, 0.0, 0);
            continue;
        }
        e->errors += mpc_one(e->sizes, e->util, e->V, e->A, &e->p, e->chunk[s], e->last_q[s], e->buffer[s], tmp, n, H,
                             mode, e->last_pred + s, e->err_ring + (size_t)s * e->K, e->err_len + s, e->K,
                             action + s, best_j ? best_j + s : 0, best_seq ? best_seq + (size_t)s * H : 0, 0, 0, 1, 0.0, 0);
    }
}

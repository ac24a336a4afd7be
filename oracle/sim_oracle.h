/* oracle/sim_oracle.h -- CPU restatement of the reference scheduling loop.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.  The product (spark-sched-sim_b200/, include/ssb.h) never links or calls it.
 *
 * A literal, scalar, single-environment restatement of
 *   spark_sched_sim/spark_sched_sim.py, components/{event,job,stage,task,executor,executor_tracker}.py,
 *   utils.py, data_samplers/tpch.py:54-106,208-235 and schedulers/heuristics/{round_robin,utils}.py
 * including CPython 3.12 `set` iteration/pop order (Objects/setobject.c) and `dict`
 * insertion order.  Pinned against traces of the unmodified reference (tests/golden/, produced by
 * oracle/refrun.py + tests/golden/gen_golden.py).
 */
#ifndef SIM_ORACLE_H
#define SIM_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
    int32_t num_executors;
    int32_t job_arrival_cap; /* <= 0: no cap (a time limit must then be given at reset) */
    double moving_delay;
    double warmup_delay;
    double job_arrival_rate;
    double beta;
} orc_config;

typedef struct {
    int32_t num_templates;
    const int32_t *num_stages;      /* [T]      */
    const int32_t *stage_base;      /* [T+1]    */
    const int32_t *edge_base;       /* [T+1]    */
    const int32_t *edges;           /* [M][2]   */
    const int32_t *num_tasks;       /* [TS]     */
    const double *rough_duration;   /* [TS]     */
    const uint8_t *present;         /* [TS][3]  */
    const uint32_t *dur_off;        /* [TS][3][8] */
    const uint32_t *dur_cnt;        /* [TS][3][8] */
    const double *dur_values;
} orc_bank;

typedef struct orc_env orc_env;

/* error codes: 0 ok; 1..9 = the reference's ValueError/KeyError paths; >= 1000 = failed assert
 * (1000 + source line of the CHECK in sim_oracle.c) */
enum {
    ORC_OK = 0,
    ORC_E_ACTION_SPACE = 1,   /* spark_sched_sim.py:276-277 */
    ORC_E_STAGE_KEY = 2,      /* :284 KeyError on stage_selection_map */
    ORC_E_NOT_SCHEDULABLE = 3,/* :286-287 */
    ORC_E_ZERO_EXEC = 4,      /* :291-292 */
    ORC_E_TOO_MANY_EXEC = 5,  /* :294-295 */
    ORC_E_NO_LIMIT = 6,       /* :137-138 */
    ORC_E_SAMPLER = 7,        /* tpch.py:106 uncaught KeyError/ValueError */
    ORC_E_TAPE_EXHAUSTED = 8,
    ORC_E_DONE = 9
};

orc_env *orc_create(const orc_config *cfg, const orc_bank *bank);
void orc_destroy(orc_env *env);

/* reset from a pre-sampled job sequence + duration tape (tape == NULL: Philox task stream with `seed`) */
int orc_reset_trace(orc_env *env, int32_t n_jobs, const double *t_arrival, const int32_t *tmpl,
                    const double *tape, int64_t n_tape, uint64_t seed);
/* reset with on-"device" sampling: Philox job + task streams (oracle/philox_ref.py) */
int orc_reset_seed(orc_env *env, uint64_t seed, double time_limit);

int orc_step(orc_env *env, int32_t stage_idx, int32_t num_exec, double *reward, int32_t *terminated);

/* observation of the last reset/step.  scalars: [N, M, Ja, num_committable, source_job_idx] */
void orc_obs_sizes(const orc_env *env, int32_t *scalars5);
void orc_obs_copy(const orc_env *env, float *nodes /*[N][3]*/, int32_t *edge_links /*[M][2]*/,
                  int32_t *dag_ptr /*[Ja+1]*/, int32_t *exec_supplies /*[Ja]*/);

/* the reference's fair/FIFO policy on the current observation (round_robin.py:14-49) */
void orc_fair_action(const orc_env *env, int32_t dynamic_partition, int32_t *stage_idx, int32_t *num_exec);

/* state queries */
double orc_wall_time(const orc_env *env);
int32_t orc_num_jobs(const orc_env *env);
int32_t orc_error(const orc_env *env);
int64_t orc_num_launches(const orc_env *env);
void orc_job_times(const orc_env *env, double *t_arrival, double *t_completed, int32_t *tmpl);

/* event log (every popped event, in pop order) */
void orc_log_enable(orc_env *env, int32_t on);
int64_t orc_log_size(const orc_env *env);
void orc_log_copy(const orc_env *env, int64_t lo, int64_t hi, double *t, uint8_t *type, int16_t *job,
                  int16_t *stage, int32_t *task, int16_t *exec, double *t_accepted);

/* executor.history (components/executor.py:25-44; only the renderer reads it): every add_history call of the episode
 * in call order -- (wall time, executor id, job id or -1 for the common pool) */
int64_t orc_history_size(const orc_env *env);
void orc_history_copy(const orc_env *env, double *t, int32_t *exec, int32_t *job);

/* whole fair/FIFO episodes back to back (CPU baseline): returns decisions made, <0 on error */
int64_t orc_run_fair_episode(orc_env *env, uint64_t seed, int32_t dynamic_partition, int64_t *events);

/* RNG spec pieces, exported for known-answer tests */
void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
double orc_neglog_u32(uint32_t w);

/* CPython-set emulation, exported so tests can fuzz it against the live interpreter */
typedef struct orc_pyset orc_pyset;
orc_pyset *orc_pyset_new(void);
void orc_pyset_free(orc_pyset *s);
void orc_pyset_add(orc_pyset *s, int32_t key);
int orc_pyset_remove(orc_pyset *s, int32_t key); /* 0 ok, -1 KeyError */
int32_t orc_pyset_pop(orc_pyset *s);             /* -1 if empty */
orc_pyset *orc_pyset_copy(const orc_pyset *s);
int32_t orc_pyset_len(const orc_pyset *s);
int32_t orc_pyset_list(const orc_pyset *s, int32_t *out); /* iteration order; returns count */

#ifdef __cplusplus
}
#endif
#endif

/* oracle/sim_oracle.c -- CPU restatement of the reference scheduling loop.  TEST INFRASTRUCTURE ONLY
 * (see sim_oracle.h).  Every function cites the reference file:line it follows; paths are relative
 * to /root/reference.  Compile with -ffp-contract=off (no FMA contraction: event times must be
 * single IEEE f64 additions, spark_sched_sim.py:610,632).
 */
#include "sim_oracle.h"

#include <math.h>
#include <setjmp.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------------
 * CPython 3.12 set of small non-negative ints (Objects/setobject.c; hash(i) == i).
 * ---------------------------------------------------------------------------------------------- */
#define PS_EMPTY (-1)
#define PS_DUMMY (-2)
#define LINEAR_PROBES 9
#define PERTURB_SHIFT 5
#define PS_MINSIZE 8

struct orc_pyset {
    int32_t mask, fill, used, finger;
    int32_t *table;
};
typedef struct orc_pyset PySet;

static void ps_init(PySet *s)
{
    s->mask = PS_MINSIZE - 1;
    s->fill = s->used = s->finger = 0;
    s->table = (int32_t *)malloc(sizeof(int32_t) * PS_MINSIZE);
    for (int i = 0; i < PS_MINSIZE; i++) s->table[i] = PS_EMPTY;
}
static void ps_clear_free(PySet *s)
{
    free(s->table);
    s->table = NULL;
}
static void ps_reinit(PySet *s)
{
    if (s->table) free(s->table);
    ps_init(s);
}

/* set_insert_clean */
static void ps_insert_clean(int32_t *table, int32_t mask, int32_t key)
{
    uint64_t perturb = (uint64_t)key;
    uint64_t i = (uint64_t)key & (uint64_t)mask;
    for (;;) {
        if (table[i] == PS_EMPTY) { table[i] = key; return; }
        if (i + LINEAR_PROBES <= (uint64_t)mask) {
            for (int j = 1; j <= LINEAR_PROBES; j++)
                if (table[i + j] == PS_EMPTY) { table[i + j] = key; return; }
        }
        perturb >>= PERTURB_SHIFT;
        i = (i * 5 + 1 + perturb) & (uint64_t)mask;
    }
}

/* set_table_resize */
static void ps_resize(PySet *s, int32_t minused)
{
    int32_t newsize = PS_MINSIZE;
    while (newsize <= minused) newsize <<= 1;
    int32_t oldmask = s->mask;
    int32_t *old = s->table;
    if (newsize == PS_MINSIZE && oldmask == PS_MINSIZE - 1 && s->fill == s->used)
        return; /* small table, no dummies: nothing to do */
    int32_t *nt = (int32_t *)malloc(sizeof(int32_t) * newsize);
    for (int i = 0; i < newsize; i++) nt[i] = PS_EMPTY;
    s->mask = newsize - 1;
    s->table = nt;
    s->fill = s->used;
    for (int i = 0; i <= oldmask; i++)
        if (old[i] >= 0) ps_insert_clean(nt, s->mask, old[i]);
    free(old);
}

/* set_add_entry */
static void ps_add(PySet *s, int32_t key)
{
    int32_t mask = s->mask;
    uint64_t perturb = (uint64_t)key;
    uint64_t i = (uint64_t)key & (uint64_t)mask;
    int64_t freeslot = -1;
    for (;;) {
        int probes = (i + LINEAR_PROBES <= (uint64_t)mask) ? LINEAR_PROBES : 0;
        for (int j = 0; j <= probes; j++) {
            int32_t v = s->table[i + j];
            if (v == PS_EMPTY) {
                if (freeslot >= 0) { /* found_unused_or_dummy: reuse the last dummy seen */
                    s->used++;
                    s->table[freeslot] = key;
                    return;
                }
                s->fill++;
                s->used++;
                s->table[i + j] = key;
                if ((int64_t)s->fill * 5 < (int64_t)mask * 3) return;
                ps_resize(s, s->used > 50000 ? s->used * 2 : s->used * 4);
                return;
            }
            if (v == key) return; /* found_active */
            if (v == PS_DUMMY) freeslot = (int64_t)(i + j);
        }
        perturb >>= PERTURB_SHIFT;
        i = (i * 5 + 1 + perturb) & (uint64_t)mask;
    }
}

/* set_lookkey: slot index of key, or -1 */
static int64_t ps_find(const PySet *s, int32_t key)
{
    int32_t mask = s->mask;
    uint64_t perturb = (uint64_t)key;
    uint64_t i = (uint64_t)key & (uint64_t)mask;
    for (;;) {
        int probes = (i + LINEAR_PROBES <= (uint64_t)mask) ? LINEAR_PROBES : 0;
        for (int j = 0; j <= probes; j++) {
            int32_t v = s->table[i + j];
            if (v == PS_EMPTY) return -1;
            if (v == key) return (int64_t)(i + j);
        }
        perturb >>= PERTURB_SHIFT;
        i = (i * 5 + 1 + perturb) & (uint64_t)mask;
    }
}

/* set_discard_entry via set.remove */
static int ps_remove(PySet *s, int32_t key)
{
    int64_t slot = ps_find(s, key);
    if (slot < 0) return -1;
    s->table[slot] = PS_DUMMY;
    s->used--;
    return 0;
}

/* set_pop */
static int32_t ps_pop(PySet *s)
{
    if (s->used == 0) return -1;
    int32_t i = s->finger & s->mask;
    while (s->table[i] < 0) {
        i++;
        if (i > s->mask) i = 0;
    }
    int32_t key = s->table[i];
    s->table[i] = PS_DUMMY;
    s->used--;
    s->finger = i + 1;
    return key;
}

/* set.copy(): make_new_set + set_merge into an empty set */
static void ps_copy_into(PySet *dst, const PySet *other)
{
    ps_reinit(dst);
    if (other->used == 0) return;
    if ((int64_t)(dst->fill + other->used) * 5 >= (int64_t)dst->mask * 3)
        ps_resize(dst, (dst->used + other->used) * 2);
    if (dst->mask == other->mask && other->fill == other->used) {
        for (int i = 0; i <= other->mask; i++) dst->table[i] = other->table[i];
        dst->fill = other->fill;
        dst->used = other->used;
        return;
    }
    dst->fill = other->used;
    dst->used = other->used;
    for (int i = 0; i <= other->mask; i++)
        if (other->table[i] >= 0) ps_insert_clean(dst->table, dst->mask, other->table[i]);
}

orc_pyset *orc_pyset_new(void)
{
    PySet *s = (PySet *)calloc(1, sizeof(PySet));
    ps_init(s);
    return s;
}
void orc_pyset_free(orc_pyset *s)
{
    ps_clear_free(s);
    free(s);
}
void orc_pyset_add(orc_pyset *s, int32_t key) { ps_add(s, key); }
int orc_pyset_remove(orc_pyset *s, int32_t key) { return ps_remove(s, key); }
int32_t orc_pyset_pop(orc_pyset *s) { return ps_pop(s); }
orc_pyset *orc_pyset_copy(const orc_pyset *s)
{
    PySet *d = (PySet *)calloc(1, sizeof(PySet));
    ps_copy_into(d, s);
    return d;
}
int32_t orc_pyset_len(const orc_pyset *s) { return s->used; }
int32_t orc_pyset_list(const orc_pyset *s, int32_t *out)
{
    int32_t n = 0;
    for (int i = 0; i <= s->mask; i++)
        if (s->table[i] >= 0) out[n++] = s->table[i];
    return n;
}

/* ------------------------------------------------------------------------------------------------
 * Philox4x32-10 + draw mapping (oracle/philox_ref.py is the spec)
 * ---------------------------------------------------------------------------------------------- */
void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; r++) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

double orc_neglog_u32(uint32_t w)
{
    uint64_t k = (uint64_t)w + 1;
    int e = 63 - __builtin_clzll(k);
    double m = (double)k / (double)((uint64_t)1 << e);
    if (m > 1.4142135623730951) { m = m * 0.5; e += 1; }
    double s = (m - 1.0) / (m + 1.0);
    double z = s * s;
    static const double D[11] = {21.0, 19.0, 17.0, 15.0, 13.0, 11.0, 9.0, 7.0, 5.0, 3.0, 1.0};
    double p = 1.0 / 23.0;
    for (int i = 0; i < 11; i++) {
        double t = p * z;
        p = t + 1.0 / D[i];
    }
    double lnm = (2.0 * s) * p;
    double el = (double)(e - 32) * 0.6931471805599453;
    return -(lnm + el);
}

static inline uint32_t bounded(uint32_t w, uint32_t n) { return (uint32_t)(((uint64_t)w * n) >> 32); }

/* ------------------------------------------------------------------------------------------------
 * Simulator state (components/*.py)
 * ---------------------------------------------------------------------------------------------- */
enum { EV_JOB_ARRIVAL = 0, EV_TASK_FINISHED = 1, EV_EXECUTOR_READY = 2 };
static const int LEVELS[8] = {5, 10, 20, 40, 50, 60, 80, 100};

typedef struct {
    int id, job_id, num_tasks, remaining, executing, completed; /* stage.py:8-18 */
    double most_recent_duration;
    int is_schedulable;
    int ts;   /* bank row */
    int node; /* all_job_ptr[job] + id */
} Stage;

typedef struct {
    int id, n_stages, tmpl;
    Stage *stages;
    int *active;  /* active_stages: stage ids in list order (job.py:22) */
    int n_active;
    uint8_t *frontier; /* frontier_stages membership (job.py:25) */
    double t_arrival, t_completed;
    uint8_t *local; /* local_executors membership */
    int n_local;
    int saturated_stage_count;
    /* dag (networkx DiGraph from the adjacency matrix, tpch.py:199) */
    int *pred_ptr, *pred, *succ_ptr, *succ;
    int n_edges;
    const int32_t *edges; /* template edges [n_edges][2] */
} Job;

typedef struct {
    int id;
    int has_task, task_stage, task_job, task_id; /* executor.task (executor.py:11) */
    int job_id;                                  /* -1 == None */
    int is_executing;
} Executor;

typedef struct {
    double t;
    int64_t counter;
    int type, job, stage, task, exec;
    double t_accepted;
} Event;

typedef struct { int src, dst, n; } Commit;

#define POOL_NONE 0
#define POOL_COMMON 1

struct orc_env {
    orc_config cfg;
    orc_bank bank;
    int E;
    double exec_intervals[256][2];
    /* per-template dag adjacency lists */
    int **t_pred_ptr, **t_pred, **t_succ_ptr, **t_succ;

    /* episode */
    int mode_tape;
    const double *tape; int64_t n_tape;
    double *tape_own;
    uint32_t key[2];
    int64_t launch_idx;
    double wall_time;
    Job *jobs; int n_jobs, jobs_cap;
    Stage *stage_store; int n_total_stages, stage_cap;
    int *all_job_ptr;
    int32_t *all_edge_links; int n_all_edges, edge_cap;
    Executor *executors;
    /* event queue (event.py) */
    Event *pq; int pq_n, pq_cap; int64_t counter;
    /* env lists */
    int *active_job_ids; int n_active;
    uint8_t *completed_job; int n_completed;
    uint8_t *selected; /* selected_stages by node */
    Stage **schedulable; int n_sched;
    Stage **tmp_sched, **tmp_sched2;
    /* tracker (executor_tracker.py) */
    int n_pools;
    PySet *pools;
    int *pool_job, *pool_stage;
    int *n_commit_from, *n_commit_to, *n_moving_to;
    int *total_exec; /* index job+1; slot 0 == None */
    int *exec_loc;
    Commit *commits; int n_commits, commits_cap;
    int source;
    /* observation of the last reset/step */
    float *obs_nodes; int32_t *obs_edges; int32_t *obs_dag_ptr; int32_t *obs_supplies;
    int obs_N, obs_M, obs_Ja, obs_ncommit, obs_src;
    uint8_t *active_stage_mask; int32_t *node_idx;
    /* log */
    int log_on; int64_t log_n, log_cap; Event *log;
    /* executor.history (executor.py:25-44): one row per add_history call, in call order */
    int64_t hist_n, hist_cap; double *hist_t; int32_t *hist_exec, *hist_job;
    int64_t n_events;
    int done;
    int error;
    jmp_buf jb;
};
typedef struct orc_env Env;

static void fail(Env *e, int code)
{
    if (!e->error) e->error = code;
    longjmp(e->jb, 1);
}
#define CHECK(cond) do { if (!(cond)) fail(e, 1000 + __LINE__); } while (0)

static inline int pool_of_job(const Env *e, int j) { (void)e; return 2 + j; }
static inline int pool_of_stage(const Env *e, const Stage *s) { return 2 + e->n_jobs + s->node; }

/* ---------------- event queue: heapq on (t, counter) (event.py:34-49) ---------------- */
static inline int ev_less(const Event *a, const Event *b)
{
    if (a->t != b->t) return a->t < b->t;
    return a->counter < b->counter;
}
static void pq_push(Env *e, Event ev)
{
    if (e->pq_n == e->pq_cap) {
        e->pq_cap = e->pq_cap ? e->pq_cap * 2 : 64;
        e->pq = (Event *)realloc(e->pq, sizeof(Event) * e->pq_cap);
    }
    ev.counter = e->counter++;
    int i = e->pq_n++;
    while (i > 0) {
        int p = (i - 1) / 2;
        if (!ev_less(&ev, &e->pq[p])) break;
        e->pq[i] = e->pq[p];
        i = p;
    }
    e->pq[i] = ev;
}
static int pq_pop(Env *e, Event *out)
{
    if (e->pq_n == 0) return 0;
    *out = e->pq[0];
    Event last = e->pq[--e->pq_n];
    int i = 0, n = e->pq_n;
    for (;;) {
        int c = 2 * i + 1;
        if (c >= n) break;
        if (c + 1 < n && ev_less(&e->pq[c + 1], &e->pq[c])) c++;
        if (!ev_less(&e->pq[c], &last)) break;
        e->pq[i] = e->pq[c];
        i = c;
    }
    if (n > 0) e->pq[i] = last;
    return 1;
}

/* ---------------- tracker (executor_tracker.py) ---------------- */
static int tracker_source_job_id(const Env *e) /* :99-103; -1 == None */
{
    if (e->source == POOL_NONE || e->source == POOL_COMMON) return -1;
    return e->pool_job[e->source];
}
static int num_committable_execs(Env *e) /* :105-111 */
{
    int n = e->pools[e->source].used - e->n_commit_from[e->source];
    CHECK(n >= 0);
    return n;
}
static Commit *find_commit(Env *e, int src, int dst)
{
    for (int i = 0; i < e->n_commits; i++)
        if (e->commits[i].src == src && e->commits[i].dst == dst) return &e->commits[i];
    return NULL;
}
static void increment_commitments(Env *e, int dst, int n) /* :224-236 */
{
    Commit *c = find_commit(e, e->source, dst);
    if (c) c->n += n;
    else {
        CHECK(e->n_commits < e->commits_cap);
        e->commits[e->n_commits].src = e->source;
        e->commits[e->n_commits].dst = dst;
        e->commits[e->n_commits].n = n;
        e->n_commits++;
    }
    e->n_commit_from[e->source] += n;
    e->n_commit_to[dst] += n;
    CHECK(e->pools[e->source].used >= e->n_commit_from[e->source]);
}
static void add_commitment(Env *e, int n, int dst) /* :146-154 */
{
    CHECK(e->source != POOL_NONE);
    int src_job = e->pool_job[e->source], dst_job = e->pool_job[dst];
    increment_commitments(e, dst, n);
    if (dst_job != src_job) e->total_exec[dst_job + 1] += n;
}
static void decrement_commitments(Env *e, int src, int dst) /* :238-249 */
{
    Commit *c = find_commit(e, src, dst);
    CHECK(c != NULL);
    c->n -= 1;
    e->n_commit_from[src] -= 1;
    e->n_commit_to[dst] -= 1;
    CHECK(e->n_commit_from[src] >= 0);
    CHECK(e->n_commit_to[dst] >= 0);
    if (c->n == 0) {
        int idx = (int)(c - e->commits);
        memmove(&e->commits[idx], &e->commits[idx + 1], sizeof(Commit) * (e->n_commits - idx - 1));
        e->n_commits--;
    }
}
static int remove_commitment(Env *e, int executor_id, int dst) /* :156-173 */
{
    int src = e->exec_loc[executor_id];
    CHECK(src != POOL_NONE);
    CHECK(find_commit(e, src, dst) != NULL); /* ValueError("no commitments from ...") */
    int src_job = e->pool_job[src], dst_job = e->pool_job[dst];
    decrement_commitments(e, src, dst);
    if (dst_job != src_job) {
        e->total_exec[dst_job + 1] -= 1;
        CHECK(e->total_exec[dst_job + 1] >= 0);
    }
    return src;
}
static int peek_commitment(Env *e, int pool) /* :175-180; POOL_NONE == None */
{
    for (int i = 0; i < e->n_commits; i++)
        if (e->commits[i].src == pool) return e->commits[i].dst;
    return POOL_NONE;
}
static void move_executor_to_pool(Env *e, int executor_id, int new_pool, int send) /* :186-220 */
{
    if (send) CHECK(new_pool != POOL_NONE && e->pool_job[new_pool] >= 0 && e->pool_stage[new_pool] >= 0);
    int old = e->exec_loc[executor_id];
    if (old != POOL_NONE) {
        CHECK(ps_remove(&e->pools[old], executor_id) == 0);
        e->exec_loc[executor_id] = POOL_NONE;
    }
    if (!send) {
        e->exec_loc[executor_id] = new_pool;
        ps_add(&e->pools[new_pool], executor_id);
        return;
    }
    e->n_moving_to[new_pool] += 1;
    int old_job = old != POOL_NONE ? e->pool_job[old] : -1;
    int new_job = e->pool_job[new_pool];
    CHECK(old_job != new_job);
    e->total_exec[new_job + 1] += 1;
    if (old_job != -1) {
        e->total_exec[old_job + 1] -= 1;
        CHECK(e->total_exec[old_job + 1] >= 0);
    }
}

/* ---------------- sampler (data_samplers/tpch.py) ---------------- */
static void init_executor_intervals(Env *e, int cap) /* tpch.py:237-262 */
{
    for (int i = 0; i <= cap; i++) e->exec_intervals[i][0] = e->exec_intervals[i][1] = 0.0;
    for (int i = 0; i <= LEVELS[0] && i <= cap; i++) e->exec_intervals[i][0] = e->exec_intervals[i][1] = LEVELS[0];
    for (int i = 0; i < 7; i++) {
        for (int r = LEVELS[i] + 1; r < LEVELS[i + 1] && r <= cap; r++) {
            e->exec_intervals[r][0] = LEVELS[i];
            e->exec_intervals[r][1] = LEVELS[i + 1];
        }
        if (LEVELS[i + 1] > cap) break;
        e->exec_intervals[LEVELS[i + 1]][0] = e->exec_intervals[LEVELS[i + 1]][1] = LEVELS[i + 1];
    }
    if (cap > LEVELS[7])
        for (int r = LEVELS[7] + 1; r < cap; r++) e->exec_intervals[r][0] = e->exec_intervals[r][1] = LEVELS[7];
}
static int level_index(double key)
{
    for (int i = 0; i < 8; i++)
        if ((double)LEVELS[i] == key) return i;
    return -1;
}
/* _sample_task_duration (tpch.py:208-214): returns 0 and sets *out, or -1 (KeyError / ValueError) */
static int sample_wave(Env *e, const Stage *st, int wave, int lvl, uint32_t w1, int warmup, double *out)
{
    const orc_bank *b = &e->bank;
    if (lvl < 0 || !((b->present[st->ts * 3 + wave] >> lvl) & 1)) return -1; /* KeyError */
    uint32_t cnt = b->dur_cnt[(st->ts * 3 + wave) * 8 + lvl];
    if (cnt == 0) return -1; /* ValueError: empty choice */
    uint32_t off = b->dur_off[(st->ts * 3 + wave) * 8 + lvl];
    double d = b->dur_values[off + bounded(w1, cnt)];
    if (warmup) d = d + e->cfg.warmup_delay;
    *out = d;
    return 0;
}
static double task_duration(Env *e, Job *job, Stage *stage, Executor *ex) /* tpch.py:75-106 */
{
    if (e->mode_tape) {
        if (e->launch_idx >= e->n_tape) fail(e, ORC_E_TAPE_EXHAUSTED);
        return e->tape[e->launch_idx];
    }
    int n_local = job->n_local;
    CHECK(n_local > 0);
    uint32_t ctr[4] = {(uint32_t)e->launch_idx, (uint32_t)((uint64_t)e->launch_idx >> 32), 2u, 0u}, w[4];
    orc_philox4x32_10(ctr, e->key, w);
    /* _sample_executor_key (tpch.py:216-235) */
    double left = e->exec_intervals[n_local][0], right = e->exec_intervals[n_local][1], key;
    if (left == right) key = left;
    else {
        double u = (double)w[0] * (1.0 / 4294967296.0);
        int rand_pt = 1 + (int)(u * (right - left));
        key = (rand_pt <= n_local - (int)left) ? left : right;
    }
    int lvl = level_index(key);
    uint8_t fw = e->bank.present[stage->ts * 3 + 1];
    if (lvl < 0 || !((fw >> lvl) & 1)) { /* key not in first_wave -> max(first_wave) */
        lvl = -1;
        for (int i = 7; i >= 0; i--) if ((fw >> i) & 1) { lvl = i; break; }
    }
    double d;
    if (!ex->has_task) { /* executor.is_idle */
        if (sample_wave(e, stage, 0, lvl, w[1], 0, &d) == 0) return d;
        if (sample_wave(e, stage, 1, lvl, w[1], 1, &d) == 0) return d;
        fail(e, ORC_E_SAMPLER);
    }
    if (ex->task_stage == stage->id) {
        if (sample_wave(e, stage, 2, lvl, w[1], 0, &d) == 0) return d;
    }
    if (sample_wave(e, stage, 1, lvl, w[1], 0, &d) == 0) return d;
    if (sample_wave(e, stage, 0, lvl, w[1], 0, &d) == 0) return d;
    fail(e, ORC_E_SAMPLER);
    return 0.0;
}

/* ---------------- env helpers (spark_sched_sim.py) ---------------- */
static int get_executor_demand(Env *e, const Stage *s) /* :566-578 */
{
    int p = pool_of_stage(e, s);
    return s->remaining - (e->n_moving_to[p] + e->n_commit_to[p]);
}
static int is_stage_saturated(Env *e, const Stage *s) { return get_executor_demand(e, s) <= 0; } /* :580-582 */
static int is_stage_ready(Env *e, const Stage *s) /* :542-555 */
{
    if (is_stage_saturated(e, s)) return 0;
    Job *job = &e->jobs[s->job_id];
    for (int k = job->pred_ptr[s->id]; k < job->pred_ptr[s->id + 1]; k++)
        if (!is_stage_saturated(e, &job->stages[job->pred[k]])) return 0;
    return 1;
}
/* :505-540.  n_ids == 0 <=> `not job_ids`; source_job_id <= 0 <=> `not source_job_id` */
static int find_schedulable_stages(Env *e, const int *job_ids, int n_ids, int source_job_id, Stage **out)
{
    if (n_ids == 0) { job_ids = e->active_job_ids; n_ids = e->n_active; }
    if (source_job_id <= 0) source_job_id = tracker_source_job_id(e);
    int n = 0;
    for (int a = 0; a < n_ids; a++) {
        int jid = job_ids[a];
        if (!(jid == source_job_id || e->total_exec[jid + 1] < e->E)) continue;
        Job *job = &e->jobs[jid];
        for (int k = 0; k < job->n_active; k++) {
            Stage *s = &job->stages[job->active[k]];
            if (!e->selected[s->node] && is_stage_ready(e, s)) out[n++] = s;
        }
    }
    return n;
}
static int job_saturated(const Job *j) { return j->saturated_stage_count == j->n_stages; } /* job.py:54-55 */
static int stage_completed(const Stage *s) { return s->completed == s->num_tasks; }        /* stage.py:38-39 */

/* Executor.add_history (executor.py:34-44): closes the executor's latest history entry with the wall time and opens
 * a new one for job_id (-1 = common pool).  Kept as the flat sequence of calls; executor k's list is
 * [[t_1, -1], [t_2, job_1], ..., [None, job_n]] for its calls (t_i, job_i). */
static void add_history(Env *e, int executor_id, int job_id)
{
    if (e->hist_n == e->hist_cap) {
        e->hist_cap = e->hist_cap ? e->hist_cap * 2 : 256;
        e->hist_t = (double *)realloc(e->hist_t, sizeof(double) * e->hist_cap);
        e->hist_exec = (int32_t *)realloc(e->hist_exec, sizeof(int32_t) * e->hist_cap);
        e->hist_job = (int32_t *)realloc(e->hist_job, sizeof(int32_t) * e->hist_cap);
    }
    e->hist_t[e->hist_n] = e->wall_time;
    e->hist_exec[e->hist_n] = executor_id;
    e->hist_job[e->hist_n] = job_id;
    e->hist_n++;
}

static void detach_executor(Env *e, Job *job, Executor *ex) /* job.py:86-89 */
{
    CHECK(job->local[ex->id]);
    job->local[ex->id] = 0;
    job->n_local--;
    ex->job_id = -1;
    ex->has_task = 0;
}
static void attach_executor(Env *e, Job *job, Executor *ex) /* job.py:81-84 */
{
    CHECK(!ex->has_task);
    if (!job->local[ex->id]) { job->local[ex->id] = 1; job->n_local++; }
    ex->job_id = job->id;
}

static void move_executor_to_stage(Env *e, Executor *ex, Stage *stage);

/* _get_idle_source_executors (:714-728): set(generator over a copy of the pool) */
static void get_idle_source_executors(Env *e, int pool /* POOL_NONE => source */, PySet *out)
{
    if (pool == POOL_NONE) pool = e->source;
    PySet cp;
    cp.table = NULL;
    ps_copy_into(&cp, &e->pools[pool]);
    ps_reinit(out);
    for (int i = 0; i <= cp.mask; i++) {
        int v = cp.table[i];
        if (v >= 0 && !e->executors[v].is_executing) ps_add(out, v);
    }
    ps_clear_free(&cp);
}

/* _move_idle_executors (:745-782).  src == POOL_NONE => None; ids == NULL => None */
static void move_idle_executors(Env *e, int src, const int *ids, int n_ids)
{
    if (src == POOL_NONE) src = e->source;
    CHECK(src != POOL_NONE);
    if (src == POOL_COMMON) return;
    int32_t buf[256];
    if (ids == NULL) {
        PySet idle;
        idle.table = NULL;
        get_idle_source_executors(e, src, &idle);
        n_ids = orc_pyset_list(&idle, buf);
        ps_clear_free(&idle);
        ids = buf;
    }
    CHECK(n_ids > 0);
    int job_id = e->pool_job[src], stage_id = e->pool_stage[src];
    CHECK(job_id >= 0);
    int sat = job_saturated(&e->jobs[job_id]);
    if (stage_id < 0 && !sat) return;
    int dst = sat ? POOL_COMMON : pool_of_job(e, job_id);
    for (int i = 0; i < n_ids; i++) {
        move_executor_to_pool(e, ids[i], dst, 0);
        if (dst == POOL_COMMON) {
            detach_executor(e, &e->jobs[job_id], &e->executors[ids[i]]);
            add_history(e, ids[i], -1); /* :782 */
        }
    }
}

static void execute_next_task(Env *e, Executor *ex, Stage *stage) /* :584-615 */
{
    CHECK(stage->remaining > 0);
    CHECK(ex->job_id == stage->job_id);
    CHECK(!ex->is_executing);
    Job *job = &e->jobs[stage->job_id];
    /* Stage.launch_next_task (stage.py:53-58): tasks pop from the end => ids descend */
    CHECK(stage->executing + stage->completed < stage->num_tasks);
    int task_id = stage->remaining - 1;
    stage->remaining -= 1;
    stage->executing += 1;
    if (stage->remaining == 0) job->saturated_stage_count += 1;
    double d = task_duration(e, job, stage, ex);
    e->launch_idx++;
    ex->has_task = 1;
    ex->task_stage = stage->id;
    ex->task_job = stage->job_id;
    ex->task_id = task_id;
    ex->is_executing = 1;
    stage->most_recent_duration = d;
    Event ev;
    ev.t = e->wall_time + d;
    ev.type = EV_TASK_FINISHED;
    ev.job = stage->job_id; ev.stage = stage->id; ev.task = task_id; ev.exec = ex->id;
    ev.t_accepted = e->wall_time;
    pq_push(e, ev);
}

static void send_executor(Env *e, Executor *ex, Stage *stage) /* :617-637 */
{
    CHECK(!ex->is_executing);
    CHECK(ex->job_id != stage->job_id);
    move_executor_to_pool(e, ex->id, pool_of_stage(e, stage), 1);
    if (ex->job_id != -1) detach_executor(e, &e->jobs[ex->job_id], ex);
    Event ev;
    ev.t = e->wall_time + e->cfg.moving_delay;
    ev.type = EV_EXECUTOR_READY;
    ev.job = stage->job_id; ev.stage = stage->id; ev.task = -1; ev.exec = ex->id;
    ev.t_accepted = INFINITY;
    pq_push(e, ev);
}

static Stage *find_backup_stage(Env *e, Executor *ex) /* :821-845 */
{
    CHECK(ex->job_id != -1);
    int own = ex->job_id;
    int n = find_schedulable_stages(e, &own, 1, own, e->tmp_sched2);
    if (n) return e->tmp_sched2[0];
    int *others = (int *)malloc(sizeof(int) * (e->n_active + 1));
    int no = 0;
    for (int i = 0; i < e->n_active; i++)
        if (e->active_job_ids[i] != ex->job_id) others[no++] = e->active_job_ids[i];
    n = find_schedulable_stages(e, others, no, own, e->tmp_sched2);
    free(others);
    if (n) return e->tmp_sched2[0];
    return NULL;
}

static void try_backup_schedule(Env *e, Executor *ex) /* :784-797 */
{
    Stage *b = find_backup_stage(e, ex);
    if (b) { move_executor_to_stage(e, ex, b); return; }
    int loc = e->exec_loc[ex->id];
    int id = ex->id;
    move_idle_executors(e, loc, &id, 1);
}

static void move_executor_to_stage(Env *e, Executor *ex, Stage *stage) /* :799-819 */
{
    if (stage->remaining == 0) { try_backup_schedule(e, ex); return; }
    if (ex->job_id != stage->job_id) { send_executor(e, ex, stage); return; }
    Job *job = &e->jobs[stage->job_id];
    if (!job->frontier[stage->id]) {
        ex->has_task = 0;
        move_executor_to_pool(e, ex->id, pool_of_job(e, stage->job_id), 0);
        return;
    }
    move_executor_to_pool(e, ex->id, pool_of_stage(e, stage), 0);
    execute_next_task(e, ex, stage);
}

static void fulfill_commitment(Env *e, int executor_id, int dst) /* :699-712 */
{
    int src = remove_commitment(e, executor_id, dst);
    if (dst == POOL_COMMON) { move_idle_executors(e, src, &executor_id, 1); return; }
    int job_id = e->pool_job[dst], stage_id = e->pool_stage[dst];
    CHECK(job_id >= 0 && stage_id >= 0);
    move_executor_to_stage(e, &e->executors[executor_id], &e->jobs[job_id].stages[stage_id]);
}

static void commit_remaining_executors(Env *e) /* :487-503 */
{
    int n = num_committable_execs(e);
    if (n > 0) add_commitment(e, n, POOL_COMMON);
}

static void fulfill_commitments_from_source(Env *e) /* :730-743 */
{
    PySet idle;
    idle.table = NULL;
    get_idle_source_executors(e, POOL_NONE, &idle);
    /* get_source_commitments(): a copy of the source's dict, insertion order */
    Commit snap[512];
    int ns = 0;
    for (int i = 0; i < e->n_commits; i++)
        if (e->commits[i].src == e->source) { CHECK(ns < 512); snap[ns++] = e->commits[i]; }
    for (int i = 0; i < ns; i++) {
        int n = snap[i].n;
        CHECK(snap[i].dst != POOL_NONE);
        CHECK(n > 0);
        while (n && idle.used) {
            int ex = ps_pop(&idle);
            fulfill_commitment(e, ex, snap[i].dst);
            n--;
        }
    }
    int left = idle.used;
    ps_clear_free(&idle);
    CHECK(left == 0);
}

static void handle_job_arrival(Env *e, Job *job) /* :428-438 */
{
    e->active_job_ids[e->n_active++] = job->id;
    int jp = pool_of_job(e, job->id);
    ps_reinit(&e->pools[jp]);
    e->n_commit_from[jp] = 0;
    e->total_exec[job->id + 1] = 0;
    for (int s = 0; s < job->n_stages; s++) {
        int sp = pool_of_stage(e, &job->stages[s]);
        ps_reinit(&e->pools[sp]);
        e->n_commit_from[sp] = 0;
        e->n_commit_to[sp] = 0;
        e->n_moving_to[sp] = 0;
    }
    if (e->pools[POOL_COMMON].used > 0) e->source = POOL_COMMON;
}

static void handle_executor_arrival(Env *e, Executor *ex, Stage *stage) /* :440-450 */
{
    Job *job = &e->jobs[stage->job_id];
    attach_executor(e, job, ex);
    add_history(e, ex->id, job->id); /* :445 */
    int sp = pool_of_stage(e, stage);
    e->n_moving_to[sp] -= 1; /* record_executor_arrival, executor_tracker.py:182-184 */
    CHECK(e->n_moving_to[sp] >= 0);
    move_executor_to_pool(e, ex->id, pool_of_job(e, job->id), 0);
    move_executor_to_stage(e, ex, stage);
}

static int record_stage_completion(Env *e, Job *job, Stage *stage) /* job.py:65-73,113-128 */
{
    int idx = -1;
    for (int k = 0; k < job->n_active; k++) if (job->active[k] == stage->id) { idx = k; break; }
    CHECK(idx >= 0);
    memmove(&job->active[idx], &job->active[idx + 1], sizeof(int) * (job->n_active - idx - 1));
    job->n_active--;
    CHECK(job->frontier[stage->id]);
    job->frontier[stage->id] = 0;
    int any = 0;
    if (!stage_completed(stage)) return 0;
    for (int k = job->succ_ptr[stage->id]; k < job->succ_ptr[stage->id + 1]; k++) {
        int c = job->succ[k];
        if (stage_completed(&job->stages[c])) continue;
        int ok = 1;
        for (int q = job->pred_ptr[c]; q < job->pred_ptr[c + 1]; q++)
            if (!stage_completed(&job->stages[job->pred[q]])) { ok = 0; break; }
        if (ok) { job->frontier[c] = 1; any = 1; }
    }
    return any;
}

static void process_job_completion(Env *e, Job *job) /* :682-697 */
{
    int jp = pool_of_job(e, job->id);
    if (e->pools[jp].used > 0) move_idle_executors(e, jp, NULL, 0);
    CHECK(e->pools[jp].used == 0);
    int idx = -1;
    for (int k = 0; k < e->n_active; k++) if (e->active_job_ids[k] == job->id) { idx = k; break; }
    CHECK(idx >= 0);
    memmove(&e->active_job_ids[idx], &e->active_job_ids[idx + 1], sizeof(int) * (e->n_active - idx - 1));
    e->n_active--;
    if (!e->completed_job[job->id]) { e->completed_job[job->id] = 1; e->n_completed++; }
    job->t_completed = e->wall_time;
}

static int handle_released_executor(Env *e, Executor *ex, Stage *stage, int frontier_changed) /* :639-660 */
{
    int dst = peek_commitment(e, pool_of_stage(e, stage));
    if (dst != POOL_NONE) { fulfill_commitment(e, ex->id, dst); return 1; }
    ex->has_task = 0;
    if (frontier_changed) {
        int id = ex->id;
        move_idle_executors(e, pool_of_stage(e, stage), &id, 1);
    }
    return 0;
}

static void handle_task_completion(Env *e, Stage *stage, int executor_id) /* :452-483 */
{
    Job *job = &e->jobs[stage->job_id];
    CHECK(executor_id >= 0);
    Executor *ex = &e->executors[executor_id];
    CHECK(!stage_completed(stage));
    stage->executing -= 1; /* stage.py:60-62 */
    stage->completed += 1;
    ex->is_executing = 0;
    if (stage->remaining > 0) { execute_next_task(e, ex, stage); return; }
    int frontier_changed = 0;
    if (stage_completed(stage)) frontier_changed = record_stage_completion(e, job, stage);
    if (job->n_active == 0) process_job_completion(e, job);
    int had = handle_released_executor(e, ex, stage, frontier_changed);
    /* _update_executor_source (:662-674) */
    if (frontier_changed) e->source = pool_of_job(e, stage->job_id);
    else if (!had) e->source = pool_of_stage(e, stage);
}

static void handle_event(Env *e, const Event *ev) /* :317-318 */
{
    e->n_events++;
    if (e->log_on) {
        if (e->log_n == e->log_cap) {
            e->log_cap = e->log_cap ? e->log_cap * 2 : 4096;
            e->log = (Event *)realloc(e->log, sizeof(Event) * e->log_cap);
        }
        e->log[e->log_n] = *ev;
        e->log[e->log_n].t = e->wall_time;
        e->log_n++;
    }
    switch (ev->type) {
    case EV_JOB_ARRIVAL: handle_job_arrival(e, &e->jobs[ev->job]); break;
    case EV_EXECUTOR_READY:
        handle_executor_arrival(e, &e->executors[ev->exec], &e->jobs[ev->job].stages[ev->stage]);
        break;
    default: handle_task_completion(e, &e->jobs[ev->job].stages[ev->stage], ev->exec); break;
    }
}

static void resume_simulation(Env *e) /* :320-343 */
{
    int n_sched = 0;
    Event ev;
    while (pq_pop(e, &ev)) {
        e->wall_time = ev.t;
        handle_event(e, &ev);
        if (!num_committable_execs(e)) continue;
        n_sched = find_schedulable_stages(e, NULL, 0, -1, e->tmp_sched);
        if (n_sched) break;
        move_idle_executors(e, POOL_NONE, NULL, 0);
        e->source = POOL_NONE;
    }
    e->n_sched = n_sched;
    memcpy(e->schedulable, e->tmp_sched, sizeof(Stage *) * n_sched);
}

static double compute_jobtime(Env *e, double wall_old, const int *old_ids, int n_old) /* :847-874 */
{
    double duration = e->wall_time - wall_old;
    if (duration == 0.0) return 0.0;
    PySet ids;
    ids.table = NULL;
    ps_reinit(&ids);
    for (int i = 0; i < n_old; i++) ps_add(&ids, old_ids[i]);
    for (int i = 0; i < e->n_active; i++) ps_add(&ids, e->active_job_ids[i]);
    double job_time = 0.0;
    double beta = e->cfg.beta;
    for (int i = 0; i <= ids.mask; i++) {
        int jid = ids.table[i];
        if (jid < 0) continue;
        Job *job = &e->jobs[jid];
        double start = job->t_arrival > wall_old ? job->t_arrival : wall_old;
        double end = job->t_completed < e->wall_time ? job->t_completed : e->wall_time;
        if (beta == 0.0) job_time += end - start;
        else job_time += exp(-beta * 1e-3 * (start - wall_old)) - exp(-beta * 1e-3 * (end - wall_old));
    }
    ps_clear_free(&ids);
    if (beta > 0.0) job_time /= beta;
    return job_time;
}

static void observe(Env *e) /* :345-406 + utils.py:5-22 */
{
    for (int i = 0; i < e->n_sched; i++) e->schedulable[i]->is_schedulable = 1;
    memset(e->active_stage_mask, 0, e->n_total_stages);
    int n = 0, src_idx = e->n_active;
    int src_job = tracker_source_job_id(e);
    e->obs_dag_ptr[0] = 0;
    for (int i = 0; i < e->n_active; i++) {
        Job *job = &e->jobs[e->active_job_ids[i]];
        if (job->id == src_job) src_idx = i;
        e->obs_supplies[i] = e->total_exec[job->id + 1];
        for (int k = 0; k < job->n_active; k++) {
            Stage *s = &job->stages[job->active[k]];
            /* np.vstack of (int, float, bool) tuples -> f64 -> astype(float32) */
            e->obs_nodes[n * 3 + 0] = (float)(double)s->remaining;
            e->obs_nodes[n * 3 + 1] = (float)s->most_recent_duration;
            e->obs_nodes[n * 3 + 2] = s->is_schedulable ? 1.0f : 0.0f;
            s->is_schedulable = 0;
            e->active_stage_mask[e->all_job_ptr[job->id] + s->id] = 1;
            n++;
        }
        e->obs_dag_ptr[i + 1] = n;
    }
    int r = 0;
    for (int v = 0; v < e->n_total_stages; v++) {
        e->node_idx[v] = 0;
        if (e->active_stage_mask[v]) e->node_idx[v] = r++;
    }
    int m = 0;
    for (int k = 0; k < e->n_all_edges; k++) {
        int u = e->all_edge_links[2 * k], v = e->all_edge_links[2 * k + 1];
        if (e->active_stage_mask[u] && e->active_stage_mask[v]) {
            e->obs_edges[2 * m] = e->node_idx[u];
            e->obs_edges[2 * m + 1] = e->node_idx[v];
            m++;
        }
    }
    e->obs_N = n;
    e->obs_M = m;
    e->obs_Ja = e->n_active;
    e->obs_ncommit = num_committable_execs(e);
    e->obs_src = src_idx;
}

static void take_action(Env *e, int stage_idx, int num_exec) /* :275-315 */
{
    if (!(stage_idx >= -1 && stage_idx < e->obs_N && num_exec >= 1 && num_exec <= e->E))
        fail(e, ORC_E_ACTION_SPACE);
    if (stage_idx == -1) { commit_remaining_executors(e); return; }
    if (stage_idx >= e->n_sched) fail(e, ORC_E_STAGE_KEY);
    Stage *stage = e->schedulable[stage_idx];
    if (num_exec == 0) fail(e, ORC_E_ZERO_EXEC);
    if (num_exec > num_committable_execs(e)) fail(e, ORC_E_TOO_MANY_EXEC);
    int demand = get_executor_demand(e, stage); /* _adjust_num_executors :557-564 */
    int n = num_exec < demand ? num_exec : demand;
    CHECK(n > 0);
    add_commitment(e, n, pool_of_stage(e, stage));
    e->selected[stage->node] = 1;
    /* bisect splice :306-315 */
    int jid = stage->job_id;
    int i = 0;
    while (i < e->n_sched && e->schedulable[i]->job_id < jid) i++; /* bisect_left */
    int hi = i + e->jobs[jid].n_active;
    if (hi > e->n_sched) hi = e->n_sched;
    int j = i;
    while (j < hi && e->schedulable[j]->job_id <= jid) j++; /* bisect_right(lo=i, hi=hi) */
    int nn = find_schedulable_stages(e, &jid, 1, -1, e->tmp_sched);
    int tail = e->n_sched - j;
    memmove(&e->schedulable[i + nn], &e->schedulable[j], sizeof(Stage *) * tail);
    memcpy(&e->schedulable[i], e->tmp_sched, sizeof(Stage *) * nn);
    e->n_sched = i + nn + tail;
}

/* ---------------- reset ---------------- */
static void free_episode(Env *e)
{
    for (int j = 0; j < e->n_jobs; j++) {
        free(e->jobs[j].active);
        free(e->jobs[j].frontier);
        free(e->jobs[j].local);
    }
    for (int p = 0; p < e->n_pools; p++) ps_clear_free(&e->pools[p]);
    free(e->pools); e->pools = NULL;
    free(e->jobs); e->jobs = NULL;
    free(e->stage_store); e->stage_store = NULL;
    free(e->all_job_ptr); e->all_job_ptr = NULL;
    free(e->all_edge_links); e->all_edge_links = NULL;
    free(e->active_job_ids); e->active_job_ids = NULL;
    free(e->completed_job); e->completed_job = NULL;
    free(e->selected); e->selected = NULL;
    free(e->schedulable); e->schedulable = NULL;
    free(e->tmp_sched); e->tmp_sched = NULL;
    free(e->tmp_sched2); e->tmp_sched2 = NULL;
    free(e->pool_job); e->pool_job = NULL;
    free(e->pool_stage); e->pool_stage = NULL;
    free(e->n_commit_from); e->n_commit_from = NULL;
    free(e->n_commit_to); e->n_commit_to = NULL;
    free(e->n_moving_to); e->n_moving_to = NULL;
    free(e->total_exec); e->total_exec = NULL;
    free(e->commits); e->commits = NULL;
    free(e->obs_nodes); e->obs_nodes = NULL;
    free(e->obs_edges); e->obs_edges = NULL;
    free(e->obs_dag_ptr); e->obs_dag_ptr = NULL;
    free(e->obs_supplies); e->obs_supplies = NULL;
    free(e->active_stage_mask); e->active_stage_mask = NULL;
    free(e->node_idx); e->node_idx = NULL;
    e->n_jobs = 0;
    e->n_pools = 0;
}

/* builds everything reset() builds once the job sequence is known (:145-186) */
static void build_episode(Env *e, int n_jobs, const double *t_arrival, const int32_t *tmpl)
{
    const orc_bank *b = &e->bank;
    free_episode(e);
    e->wall_time = 0;
    e->pq_n = 0;
    e->counter = 0;
    e->launch_idx = 0;
    e->log_n = 0;
    e->hist_n = 0;
    e->n_events = 0;
    e->done = 0;
    e->n_jobs = n_jobs;
    e->jobs = (Job *)calloc(n_jobs, sizeof(Job));
    e->all_job_ptr = (int *)calloc(n_jobs + 1, sizeof(int));
    int S = 0, M = 0;
    for (int j = 0; j < n_jobs; j++) { S += b->num_stages[tmpl[j]]; M += b->edge_base[tmpl[j] + 1] - b->edge_base[tmpl[j]]; }
    e->n_total_stages = S;
    e->stage_store = (Stage *)calloc(S, sizeof(Stage));
    e->all_edge_links = (int32_t *)calloc(2 * (M + 1), sizeof(int32_t));
    e->n_all_edges = M;
    int sb = 0, eb = 0;
    for (int j = 0; j < n_jobs; j++) {
        Job *job = &e->jobs[j];
        int t = tmpl[j];
        job->id = j;
        job->tmpl = t;
        job->n_stages = b->num_stages[t];
        job->stages = &e->stage_store[sb];
        job->active = (int *)calloc(job->n_stages, sizeof(int));
        job->frontier = (uint8_t *)calloc(job->n_stages, 1);
        job->local = (uint8_t *)calloc(e->E, 1);
        job->n_active = job->n_stages;
        job->t_arrival = t_arrival[j];
        job->t_completed = INFINITY;
        job->pred_ptr = e->t_pred_ptr[t]; job->pred = e->t_pred[t];
        job->succ_ptr = e->t_succ_ptr[t]; job->succ = e->t_succ[t];
        job->n_edges = b->edge_base[t + 1] - b->edge_base[t];
        job->edges = &b->edges[2 * b->edge_base[t]];
        e->all_job_ptr[j] = sb;
        for (int s = 0; s < job->n_stages; s++) {
            Stage *st = &job->stages[s];
            st->id = s; st->job_id = j;
            st->ts = b->stage_base[t] + s;
            st->num_tasks = st->remaining = b->num_tasks[st->ts];
            st->most_recent_duration = b->rough_duration[st->ts];
            st->node = sb + s;
            job->active[s] = s;
            /* _init_frontier (job.py:93-111): in-degree 0 */
            job->frontier[s] = (job->pred_ptr[s + 1] == job->pred_ptr[s]);
        }
        /* _reset_edge_links (:249-258) */
        for (int k = 0; k < job->n_edges; k++) {
            e->all_edge_links[2 * (eb + k)] = sb + job->edges[2 * k];
            e->all_edge_links[2 * (eb + k) + 1] = sb + job->edges[2 * k + 1];
        }
        sb += job->n_stages;
        eb += job->n_edges;
    }
    e->all_job_ptr[n_jobs] = sb;
    /* all arrivals pushed up-front (:152-154) */
    for (int j = 0; j < n_jobs; j++) {
        Event ev;
        ev.t = t_arrival[j]; ev.type = EV_JOB_ARRIVAL; ev.job = j; ev.stage = -1; ev.task = -1; ev.exec = -1;
        ev.t_accepted = INFINITY;
        pq_push(e, ev);
    }
    /* executors + tracker.reset (executor_tracker.py:32-70) */
    for (int x = 0; x < e->E; x++) {
        Executor *ex = &e->executors[x];
        ex->id = x; ex->has_task = 0; ex->job_id = -1; ex->is_executing = 0;
        ex->task_stage = ex->task_job = ex->task_id = -1;
    }
    e->n_pools = 2 + n_jobs + S;
    e->pools = (PySet *)calloc(e->n_pools, sizeof(PySet));
    for (int p = 0; p < e->n_pools; p++) ps_init(&e->pools[p]);
    e->pool_job = (int *)calloc(e->n_pools, sizeof(int));
    e->pool_stage = (int *)calloc(e->n_pools, sizeof(int));
    e->pool_job[0] = e->pool_job[1] = -1;
    e->pool_stage[0] = e->pool_stage[1] = -1;
    for (int j = 0; j < n_jobs; j++) {
        e->pool_job[2 + j] = j; e->pool_stage[2 + j] = -1;
        for (int s = 0; s < e->jobs[j].n_stages; s++) {
            int p = 2 + n_jobs + e->jobs[j].stages[s].node;
            e->pool_job[p] = j; e->pool_stage[p] = s;
        }
    }
    e->n_commit_from = (int *)calloc(e->n_pools, sizeof(int));
    e->n_commit_to = (int *)calloc(e->n_pools, sizeof(int));
    e->n_moving_to = (int *)calloc(e->n_pools, sizeof(int));
    e->total_exec = (int *)calloc(n_jobs + 1, sizeof(int));
    e->commits_cap = 4 * e->E + 16;
    e->commits = (Commit *)calloc(e->commits_cap, sizeof(Commit));
    e->n_commits = 0;
    for (int x = 0; x < e->E; x++) { ps_add(&e->pools[POOL_COMMON], x); e->exec_loc[x] = POOL_COMMON; }
    e->source = POOL_COMMON;
    e->active_job_ids = (int *)calloc(n_jobs + 1, sizeof(int));
    e->n_active = 0;
    e->completed_job = (uint8_t *)calloc(n_jobs + 1, 1);
    e->n_completed = 0;
    e->selected = (uint8_t *)calloc(S + 1, 1);
    e->schedulable = (Stage **)calloc(S + 1, sizeof(Stage *));
    e->tmp_sched = (Stage **)calloc(S + 1, sizeof(Stage *));
    e->tmp_sched2 = (Stage **)calloc(S + 1, sizeof(Stage *));
    e->n_sched = 0;
    e->obs_nodes = (float *)calloc(3 * (S + 1), sizeof(float));
    e->obs_edges = (int32_t *)calloc(2 * (M + 1), sizeof(int32_t));
    e->obs_dag_ptr = (int32_t *)calloc(n_jobs + 2, sizeof(int32_t));
    e->obs_supplies = (int32_t *)calloc(n_jobs + 1, sizeof(int32_t));
    e->active_stage_mask = (uint8_t *)calloc(S + 1, 1);
    e->node_idx = (int32_t *)calloc(S + 1, sizeof(int32_t));
    /* _load_initial_jobs (:260-273) */
    while (e->pq_n) {
        if (e->pq[0].t > 0) break;
        Event ev;
        pq_pop(e, &ev);
        handle_job_arrival(e, &e->jobs[ev.job]);
    }
    e->n_sched = find_schedulable_stages(e, NULL, 0, -1, e->schedulable);
    observe(e);
}

orc_env *orc_create(const orc_config *cfg, const orc_bank *bank)
{
    if (cfg->num_executors < 1 || cfg->num_executors > 255) return NULL;
    Env *e = (Env *)calloc(1, sizeof(Env));
    e->cfg = *cfg;
    e->bank = *bank;
    e->E = cfg->num_executors;
    init_executor_intervals(e, e->E);
    e->executors = (Executor *)calloc(e->E, sizeof(Executor));
    e->exec_loc = (int *)calloc(e->E, sizeof(int));
    int T = bank->num_templates;
    e->t_pred_ptr = (int **)calloc(T, sizeof(int *));
    e->t_pred = (int **)calloc(T, sizeof(int *));
    e->t_succ_ptr = (int **)calloc(T, sizeof(int *));
    e->t_succ = (int **)calloc(T, sizeof(int *));
    for (int t = 0; t < T; t++) {
        int n = bank->num_stages[t];
        int m = bank->edge_base[t + 1] - bank->edge_base[t];
        const int32_t *ed = &bank->edges[2 * bank->edge_base[t]];
        int *pp = (int *)calloc(n + 1, sizeof(int)), *sp = (int *)calloc(n + 1, sizeof(int));
        int *pr = (int *)calloc(m + 1, sizeof(int)), *su = (int *)calloc(m + 1, sizeof(int));
        for (int k = 0; k < m; k++) { sp[ed[2 * k] + 1]++; pp[ed[2 * k + 1] + 1]++; }
        for (int s = 0; s < n; s++) { sp[s + 1] += sp[s]; pp[s + 1] += pp[s]; }
        int *pc = (int *)calloc(n + 1, sizeof(int)), *sc = (int *)calloc(n + 1, sizeof(int));
        for (int k = 0; k < m; k++) { /* edges are in (u, v) row-major order => adjacency in insertion order */
            int u = ed[2 * k], v = ed[2 * k + 1];
            su[sp[u] + sc[u]++] = v;
            pr[pp[v] + pc[v]++] = u;
        }
        free(pc); free(sc);
        e->t_pred_ptr[t] = pp; e->t_pred[t] = pr; e->t_succ_ptr[t] = sp; e->t_succ[t] = su;
    }
    return e;
}

void orc_destroy(orc_env *e)
{
    if (!e) return;
    free_episode(e);
    for (int t = 0; t < e->bank.num_templates; t++) {
        free(e->t_pred_ptr[t]); free(e->t_pred[t]); free(e->t_succ_ptr[t]); free(e->t_succ[t]);
    }
    free(e->t_pred_ptr); free(e->t_pred); free(e->t_succ_ptr); free(e->t_succ);
    free(e->executors); free(e->exec_loc); free(e->pq); free(e->log); free(e->tape_own);
    free(e->hist_t); free(e->hist_exec); free(e->hist_job);
    free(e);
}

int orc_reset_trace(orc_env *e, int32_t n_jobs, const double *t_arrival, const int32_t *tmpl,
                    const double *tape, int64_t n_tape, uint64_t seed)
{
    e->error = 0;
    if (setjmp(e->jb)) return e->error;
    free(e->tape_own);
    e->tape_own = NULL;
    e->mode_tape = tape != NULL;
    if (tape) {
        e->tape_own = (double *)malloc(sizeof(double) * (n_tape + 1));
        memcpy(e->tape_own, tape, sizeof(double) * n_tape);
    }
    e->tape = e->tape_own;
    e->n_tape = n_tape;
    e->key[0] = (uint32_t)seed;
    e->key[1] = (uint32_t)(seed >> 32);
    if (n_jobs <= 0 || t_arrival[0] != 0) fail(e, 1000 + __LINE__); /* :150 first job must arrive at t=0 */
    build_episode(e, n_jobs, t_arrival, tmpl);
    return 0;
}

int orc_reset_seed(orc_env *e, uint64_t seed, double time_limit)
{
    e->error = 0;
    if (setjmp(e->jb)) return e->error;
    int cap = e->cfg.job_arrival_cap;
    if (isinf(time_limit) && cap <= 0) fail(e, ORC_E_NO_LIMIT); /* :137-138 */
    e->mode_tape = 0;
    e->key[0] = (uint32_t)seed;
    e->key[1] = (uint32_t)(seed >> 32);
    /* job_sequence (tpch.py:54-73) on the Philox job stream */
    int n = 0, capn = 64;
    double *ta = (double *)malloc(sizeof(double) * capn);
    int32_t *tm = (int32_t *)malloc(sizeof(int32_t) * capn);
    double t = 0;
    double mean = 1 / e->cfg.job_arrival_rate;
    while (t < time_limit && (cap <= 0 || n < cap)) {
        if (n == capn) {
            capn *= 2;
            ta = (double *)realloc(ta, sizeof(double) * capn);
            tm = (int32_t *)realloc(tm, sizeof(int32_t) * capn);
        }
        uint32_t ctr[4] = {(uint32_t)n, 0u, 1u, 0u}, w[4];
        orc_philox4x32_10(ctr, e->key, w);
        int query = (int)bounded(w[0], 22);           /* 1 + integers(22), tpch.py:177 */
        int size = (int)bounded(w[1], 7);             /* choice(QUERY_SIZES), tpch.py:178 */
        ta[n] = t;
        tm[n] = size * 22 + query;
        n++;
        double x = mean * orc_neglog_u32(w[2]);      /* exponential(mean), tpch.py:70 */
        t = t + x;
    }
    if (n == 0) { free(ta); free(tm); fail(e, 1000 + __LINE__); }
    /* build_episode may longjmp; ta/tm are small and leak only on failure */
    build_episode(e, n, ta, tm);
    free(ta);
    free(tm);
    return 0;
}

int orc_step(orc_env *e, int32_t stage_idx, int32_t num_exec, double *reward, int32_t *terminated)
{
    if (e->error) return e->error;
    if (setjmp(e->jb)) return e->error;
    if (e->done) fail(e, ORC_E_DONE);
    *reward = 0;
    *terminated = 0;
    take_action(e, stage_idx, num_exec);
    if (num_committable_execs(e) && e->n_sched) { /* :191-193 */
        observe(e);
        return 0;
    }
    commit_remaining_executors(e);
    fulfill_commitments_from_source(e);
    e->source = POOL_NONE;
    for (int i = 0; i < e->n_total_stages; i++) e->selected[i] = 0;
    double wall_old = e->wall_time;
    int n_old = e->n_active;
    int *old_ids = (int *)malloc(sizeof(int) * (n_old + 1));
    memcpy(old_ids, e->active_job_ids, sizeof(int) * n_old);
    resume_simulation(e);
    double jt = compute_jobtime(e, wall_old, old_ids, n_old);
    free(old_ids);
    *reward = -jt;
    int term = e->n_completed == e->n_jobs;
    *terminated = term;
    if (!term) CHECK(num_committable_execs(e) && e->n_sched); /* :212-215 */
    else e->done = 1;
    observe(e);
    return 0;
}

void orc_obs_sizes(const orc_env *e, int32_t *sc)
{
    sc[0] = e->obs_N; sc[1] = e->obs_M; sc[2] = e->obs_Ja; sc[3] = e->obs_ncommit; sc[4] = e->obs_src;
}
void orc_obs_copy(const orc_env *e, float *nodes, int32_t *edge_links, int32_t *dag_ptr, int32_t *sup)
{
    memcpy(nodes, e->obs_nodes, sizeof(float) * 3 * e->obs_N);
    memcpy(edge_links, e->obs_edges, sizeof(int32_t) * 2 * e->obs_M);
    memcpy(dag_ptr, e->obs_dag_ptr, sizeof(int32_t) * (e->obs_Ja + 1));
    memcpy(sup, e->obs_supplies, sizeof(int32_t) * e->obs_Ja);
}

/* heuristics/utils.py:17-37 `find_stage` on the stored observation */
static int fair_find_stage(const Env *e, const uint8_t *frontier, const int32_t *rank, int job_idx)
{
    int sel = -1;
    for (int node = e->obs_dag_ptr[job_idx]; node < e->obs_dag_ptr[job_idx + 1]; node++) {
        if (rank[node] < 0) continue;
        if (frontier[node]) return rank[node];
        if (sel == -1) sel = rank[node];
    }
    return sel;
}

void orc_fair_action(const orc_env *e, int32_t dynamic_partition, int32_t *stage_idx, int32_t *num_exec)
{
    /* preprocess_obs (heuristics/utils.py:5-14) */
    int N = e->obs_N;
    uint8_t *frontier = (uint8_t *)malloc(N + 1);
    int32_t *rank = (int32_t *)malloc(sizeof(int32_t) * (N + 1));
    memset(frontier, 1, N + 1);
    for (int k = 0; k < e->obs_M; k++) frontier[e->obs_edges[2 * k + 1]] = 0;
    int r = 0;
    for (int v = 0; v < N; v++) rank[v] = (e->obs_nodes[3 * v + 2] != 0.0f) ? r++ : -1;
    /* RoundRobinScheduler.schedule (round_robin.py:14-49) */
    int Ja = e->obs_Ja;
    int cap = dynamic_partition ? (e->E + (Ja > 1 ? Ja : 1) - 1) / (Ja > 1 ? Ja : 1) : e->E;
    *stage_idx = -1;
    *num_exec = e->obs_ncommit;
    if (e->obs_src < Ja) {
        int sel = fair_find_stage(e, frontier, rank, e->obs_src);
        if (sel != -1) { *stage_idx = sel; goto out; }
    }
    for (int j = 0; j < Ja; j++) {
        if (e->obs_supplies[j] >= cap || j == e->obs_src) continue;
        int sel = fair_find_stage(e, frontier, rank, j);
        if (sel == -1) continue;
        int n = cap - e->obs_supplies[j];
        *stage_idx = sel;
        *num_exec = e->obs_ncommit < n ? e->obs_ncommit : n;
        goto out;
    }
out:
    free(frontier);
    free(rank);
}

double orc_wall_time(const orc_env *e) { return e->wall_time; }
int32_t orc_num_jobs(const orc_env *e) { return e->n_jobs; }
int32_t orc_error(const orc_env *e) { return e->error; }
int64_t orc_num_launches(const orc_env *e) { return e->launch_idx; }
void orc_job_times(const orc_env *e, double *ta, double *tc, int32_t *tmpl)
{
    for (int j = 0; j < e->n_jobs; j++) {
        ta[j] = e->jobs[j].t_arrival;
        tc[j] = e->jobs[j].t_completed;
        if (tmpl) tmpl[j] = e->jobs[j].tmpl;
    }
}
void orc_log_enable(orc_env *e, int32_t on) { e->log_on = on; }
int64_t orc_log_size(const orc_env *e) { return e->log_n; }
void orc_log_copy(const orc_env *e, int64_t lo, int64_t hi, double *t, uint8_t *type, int16_t *job,
                  int16_t *stage, int32_t *task, int16_t *exec, double *tacc)
{
    for (int64_t i = lo; i < hi; i++) {
        const Event *ev = &e->log[i];
        int64_t k = i - lo;
        t[k] = ev->t; type[k] = (uint8_t)ev->type; job[k] = (int16_t)ev->job; stage[k] = (int16_t)ev->stage;
        task[k] = ev->task; exec[k] = (int16_t)ev->exec; tacc[k] = ev->t_accepted;
    }
}

int64_t orc_history_size(const orc_env *e) { return e->hist_n; }
void orc_history_copy(const orc_env *e, double *t, int32_t *exec, int32_t *job)
{
    for (int64_t i = 0; i < e->hist_n; i++) { t[i] = e->hist_t[i]; exec[i] = e->hist_exec[i]; job[i] = e->hist_job[i]; }
}

int64_t orc_run_fair_episode(orc_env *e, uint64_t seed, int32_t dynamic_partition, int64_t *events)
{
    int rc = orc_reset_seed(e, seed, INFINITY);
    if (rc) return -rc;
    int64_t steps = 0;
    int32_t term = 0;
    double reward;
    while (!term) {
        int32_t a, n;
        orc_fair_action(e, dynamic_partition, &a, &n);
        rc = orc_step(e, a, n, &reward, &term);
        if (rc) return -rc;
        steps++;
    }
    if (events) *events = e->n_events;
    return steps;
}

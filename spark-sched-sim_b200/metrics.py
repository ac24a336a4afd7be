"""Episode metrics with the reference's definitions (spark_sched_sim/metrics.py:4-23).  `env` is
anything exposing `.unwrapped` with `jobs`, `active_job_ids`, `completed_job_ids`, `wall_time`
(the single-env facade), so `metrics.avg_job_duration(env)` reads as in examples.py:97."""
import numpy as np


def job_durations(env):
    """Time in system of every job that has arrived; unfinished jobs are clipped at wall_time."""
    u = env.unwrapped
    ids = list(u.active_job_ids) + list(u.completed_job_ids)
    arrival = np.array([u.jobs[j].t_arrival for j in ids], dtype=np.float64)
    end = np.minimum(np.array([u.jobs[j].t_completed for j in ids], dtype=np.float64), u.wall_time)
    return list(end - arrival)


def avg_job_duration(env):
    return np.mean(job_durations(env))


def avg_num_jobs(env):
    # Little's law estimate: total job-time divided by elapsed time
    return sum(job_durations(env)) / env.unwrapped.wall_time


def job_duration_percentiles(env):
    return np.percentile(job_durations(env), [25, 50, 75, 100])

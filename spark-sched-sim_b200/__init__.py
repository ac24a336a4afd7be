"""spark-sched-sim, batched on B200: the reference's discrete-event scheduling loop
(`SparkSchedSimEnv.reset/step`, spark_sched_sim/spark_sched_sim.py) as hand-written sm_100a CUDA
kernels behind a C ABI (include/ssb.h), with a Python host layer that mirrors the reference's
Gymnasium env, scheduler and wrapper interfaces."""

try:  # the reference registers its env with gymnasium (spark_sched_sim/__init__.py:3-6); do the same when it is there
    from gymnasium.envs.registration import register as _register

    _register(id="SparkSchedSimEnv-v0", entry_point="spark_sched_sim_b200.env:SparkSchedSimEnv")
except Exception:  # gymnasium absent (this image) or the id already registered
    pass

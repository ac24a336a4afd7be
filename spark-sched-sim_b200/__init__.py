"""spark-sched-sim, batched on B200: the reference's discrete-event scheduling loop
(`SparkSchedSimEnv.reset/step`, spark_sched_sim/spark_sched_sim.py) as hand-written sm_100a CUDA
kernels behind a C ABI (include/ssb.h), with a Python host layer that mirrors the reference's
Gymnasium env, scheduler and wrapper interfaces."""

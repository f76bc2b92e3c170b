import os, sys, time, numpy as np
sys.path.insert(0, os.getcwd())
from bench import load_workload
from grid2op_b200.rollout import BatchedDoNothing
gm, chron = load_workload()
env = BatchedDoNothing(gm, chron, 4096)
for _ in range(5): env.step_host()
N=200
t_take=t_run=t_read=0.0
for _ in range(N):
    t0=time.perf_counter()
    np.take(env._chron_flat, env._row_base + env._t_host, axis=0, out=env._rows[:env.batch]); env._t_host=(env._t_host+1)%chron.shape[1]
    t1=time.perf_counter()
    env.engine.run_rows_staged(env.batch, nb_cap=env.nb_cap)
    t2=time.perf_counter()
    m=float(env._stage["out"][:env.batch][:, 3*gm.n_line:4*gm.n_line].max())
    t3=time.perf_counter()
    t_take+=t1-t0; t_run+=t2-t1; t_read+=t3-t2
print("per step us: take %.1f run_rows_staged %.1f host read %.1f"%(1e6*t_take/N,1e6*t_run/N,1e6*t_read/N))

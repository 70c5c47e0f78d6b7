"""Multi-process path: contiguous block sharding of the problem batch, no hot-path collective, one final
gather of the outcome codes.  Two gloo ranks on CPU run the closed loop on their shards with the oracle as the compute
backend (the CUDA engine needs a GPU; the sharding / gathering code under test is backend-agnostic) and must reproduce
the single-process result problem by problem."""
import os
import subprocess
import sys

import numpy as np

from safe_mpc_b200 import distributed as D

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..')

WORKER = r'''
import os, sys, json
import numpy as np
sys.path.insert(0, sys.argv[1])
import bench
from safe_mpc_b200 import distributed as D
from oracle.oracle import Oracle, OracleSim
rank, local_rank, world = D.init('gloo')
total, N, steps = 7, 8, 3
lo, hi = D.shard_range(total, rank, world)
params, md, x0, pin = bench.workload('htwa', N, 5.0, 3, lo, hi)
main, bk, prob = bench.make_handles(Oracle, params, md, 'htwa', hi - lo, 1)
main.set_plant_inertial(pin)
bench.warm_guess(main, x0, N, 2)
sim = OracleSim(main, bk, steps); sim.reset(x0); sim.run(steps)
x_log, u_log = sim.log()
D.barrier()
outcome = D.gather_outcomes(sim.outcome())
sizes = D.all_gather_vector([hi - lo, float(np.nansum(u_log))])
if rank == 0:
    print('RESULT ' + json.dumps({'outcome': outcome.tolist(), 'sizes': sizes}))
'''


def _run(world):
    procs = []
    port = 29600 + os.getpid() % 300 + world
    for r in range(world):
        env = dict(os.environ, RANK=str(r), LOCAL_RANK=str(r), WORLD_SIZE=str(world), MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port),
                   OMP_NUM_THREADS='1')
        procs.append(subprocess.Popen([sys.executable, '-c', WORKER, ROOT], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True))
    outs = [p.communicate(timeout=300) for p in procs]
    for p, (o, e) in zip(procs, outs):
        assert p.returncode == 0, e[-2000:]
    import json
    line = [l for l in outs[0][0].splitlines() if l.startswith('RESULT ')][0]
    return json.loads(line[7:])


def test_shard_range_covers_everything_once():
    for total in (1, 7, 100, 10007):
        for world in (1, 2, 3, 8):
            spans = [D.shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def test_two_gloo_ranks_reproduce_the_single_process_result():
    one = _run(1)
    two = _run(2)
    assert two['outcome'] == one['outcome'] and len(two['outcome']) == 7
    assert [int(s[0]) for s in two['sizes']] == [4, 3]
    # same controls problem by problem: the per-problem input streams do not depend on the sharding
    assert abs(sum(s[1] for s in two['sizes']) - one['sizes'][0][1]) <= 1e-9 * max(1.0, abs(one['sizes'][0][1]))


def test_outcome_counts_partition():
    o = np.array([0, 1, 2, 4, 5, 6, 4, 0])
    c = D.outcome_counts(o)
    assert c == {'completed': 2, 'collisions': 2, 'viable': 2, 'not_converged': 2}
    assert sum(c.values()) == len(o)

"""CPU suite: the N > 1 path (instance sharding + whole-job aggregation) with gloo, world_size 2."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys
sys.path.insert(0, %r)
import torch, torch.distributed as dist
from warm_start_hmpc_b200.closed_loop import shard, reduce_stats
dist.init_process_group('gloo')
r, w = dist.get_rank(), dist.get_world_size()
lo, hi = shard(1001, r, w)
units, t = reduce_stats(hi - lo, 10. * (r + 1))
assert units == 1001, units
assert t == 10. * w, t
if r == 0:
    print('OK', units, t)
dist.destroy_process_group()
'''


def test_gloo_world2_shard_and_reduce(tmp_path):
    f = tmp_path / 'worker.py'
    f.write_text(WORKER % ROOT)
    out = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node=2',
                          '--master-addr', '127.0.0.1', '--master-port', '29533', str(f)],
                         capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    assert 'OK 1001 20.0' in out.stdout

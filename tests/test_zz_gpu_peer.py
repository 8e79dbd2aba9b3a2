"""The exchange over peer memory (laml_b200/csrc/pmm_peer.cuh: count reduction + reduce-scatter + all-gather in the
E-step's finishing kernel) on ONE GPU: the ranks are handles of this process (``pmm_peer_setup_local``), each with its
own stream and a site shard, so that the kernels of all ranks are in flight together and synchronise through their
flags exactly as ranks on different GPUs do.  Checked: every rank ends with the same bits; they equal the sum of the
ranks' own statistics taken in rank order; they agree with one engine holding all sites; the EM loop on shards walks
the single engine's iterations (tests/peer_local_worker.py).  (Two GPUs, two processes, CUDA IPC: tests/test_zz_gpu_nccl.py.)

The worker runs in its own process with 32 hardware work queues: kernels of different handles that wait for each other
must not share a queue (see the worker's header), and with a patience of 2 s instead of the default 30.  If a wait
nevertheless gives up (exit code 3: a failed call, by design, instead of a hung GPU) the case is repeated; any wrong
NUMBER fails at once."""
import os
import subprocess
import sys

import pytest

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(600)]
HERE = os.path.dirname(os.path.abspath(__file__))




@pytest.mark.parametrize("world,n,K,chunks", [(2, 700, 300, 0), (3, 90, 50, 0), (4, 2000, 520, 0), (8, 300, 64, 0), (2, 500, 400, 3)])
def test_ranks_in_one_process(world, n, K, chunks):
    """chunks > 0: every rank walks its sites in that many chunks (PMM_CHUNKS: what a rank does when its records do not
    fit the memory budget) -- the earlier chunks' sums accumulate in the rank's own buffer, the last chunk's finishing
    kernel adds its part and exchanges"""
    env = dict(os.environ, CUDA_DEVICE_MAX_CONNECTIONS="32", PMM_SLICES="1", PMM_PEER_TIMEOUT_MS="2000")
    if chunks:
        env["PMM_CHUNKS"] = str(chunks)
    cmd = [sys.executable, os.path.join(HERE, "peer_local_worker.py"), str(world), str(n), str(K)]
    for attempt in range(4):
        res = subprocess.run(cmd, capture_output=True, text=True, timeout=500, env=env)
        if res.returncode != 3:
            break
    assert res.returncode == 0, "attempt %d, rc %d\n%s\n%s" % (attempt + 1, res.returncode, res.stdout[-2000:], res.stderr[-4000:])

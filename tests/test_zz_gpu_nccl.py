"""The multi-GPU data plane on hardware (needs >= 2 CUDA devices; skipped otherwise): two ranks under
``torch.distributed.run`` with the NCCL backend against one engine holding the whole problem -- see tests/nccl_worker.py.
The same host logic runs on CPU with ``gloo`` stand-ins in tests/test_sharded_gloo.py."""
import json
import os
import socket
import subprocess
import sys

import pytest

pytestmark = [pytest.mark.gpu, pytest.mark.multigpu, pytest.mark.timeout(900)]
HERE = os.path.dirname(os.path.abspath(__file__))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("world,peer", [(2, "1"), (2, "0")])
def test_site_shards_over_nccl_match_one_gpu(world, peer, tmp_path):
    """peer = "1": the exchange inside the engine's finishing kernel over peer memory (the default); "0": NCCL all-reduce"""
    out = tmp_path / "nccl.json"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(HERE, "nccl_worker.py"), str(out)]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=800, env=dict(os.environ, PMM_PEER=peer))
    assert res.returncode == 0, res.stderr[-4000:]
    ranks = json.load(open(out))
    assert len(ranks) == world
    for r in ranks:
        assert r["peer"] == (peer == "1"), r
        assert r["llh_rel"] <= 1e-12 and r["S_max_err"] <= 1.0 and r["Rtilde_rel"] <= 1e-12, r
        assert r["em_iters"][0] == r["em_iters"][1] and r["em_nllh_rel"] <= 1e-10, r
        assert r["em_d_max_abs"] <= 1e-7 and r["em_nu_abs"] <= 1e-7 and r["em_phi_abs"] <= 1e-7, r
        assert r["em_step_d_max_abs"] <= 1e-9 and r["em_step_llh_rel"] <= 1e-12, r
        assert r["ultra_iters"][0] == r["ultra_iters"][1] and r["ultra_status"][0] == r["ultra_status"][1] == "optimal", r
        assert r["ultra_nllh_rel"] <= 1e-9 and r["ultra_d_max_abs"] <= 1e-6, r
        assert r["ranks_bit_identical"] and r["nni_scores_equal"], r

"""`python -m infinistore.server` as a subprocess + the HTTP manage plane, on CPU."""
import json
import os
import signal
import subprocess
import sys
import time
import urllib.error
import urllib.parse
import urllib.request

import pytest
import torch

from conftest import ROOT, free_port, make_conn


CKPT_DIR = f"/tmp/istore_cli_ckpt_{os.getpid()}"


@pytest.fixture(scope="module")
def cli_server():
    sport, mport = free_port(), free_port()
    os.makedirs(CKPT_DIR, exist_ok=True)
    env = dict(os.environ, PYTHONPATH=ROOT + os.pathsep + os.environ.get("PYTHONPATH", ""))
    proc = subprocess.Popen(
        [sys.executable, "-m", "infinistore.server", "--service-port", str(sport),
         "--manage-port", str(mport), "--host", "127.0.0.1", "--pool-backend", "host",
         "--prealloc-size", "1", "--minimal-allocate-size", "16", "--log-level", "warning",
         "--dev-name", "mlx5_2", "--link-type", "Ethernet", "--checkpoint-dir", CKPT_DIR],
        env=env, cwd=ROOT)
    deadline = time.time() + 60
    while time.time() < deadline:
        try:
            urllib.request.urlopen(f"http://127.0.0.1:{mport}/kvmap_len", timeout=1).read()
            break
        except Exception:
            if proc.poll() is not None:
                raise RuntimeError("server exited early")
            time.sleep(0.2)
    yield sport, mport
    os.kill(proc.pid, signal.SIGINT)
    try:
        proc.wait(20)
    except subprocess.TimeoutExpired:
        proc.kill()


def _http(method, url):
    req = urllib.request.Request(url, method=method)
    return json.loads(urllib.request.urlopen(req, timeout=30).read())


def test_manage_plane_and_selftest(cli_server):
    sport, mport = cli_server
    base = f"http://127.0.0.1:{mport}"
    assert _http("GET", base + "/kvmap_len") == {"len": 0}
    conn = make_conn(sport)
    src = torch.randn(8 * 1024)
    conn.register_mr(src)
    keys = [f"cli-{i}" for i in range(8)]
    conn.rdma_write_cache(src, [i * 1024 for i in range(8)], 1024, conn.allocate_rdma(keys, 4096))
    conn.sync()
    assert _http("GET", base + "/kvmap_len") == {"len": 8}
    assert _http("POST", base + f"/selftest/{sport}") == {"status": "ok"}
    stats = _http("GET", base + "/stats")
    assert stats["keys"] == 11 and stats["ops"]["ALLOCATE"] >= 2
    text = urllib.request.urlopen(base + "/metrics", timeout=10).read().decode()
    assert "infinistore_keys 11" in text and 'infinistore_op_total{op="COMMIT"}' in text
    assert 'infinistore_op_total{op="STAGE_COMMIT"}' in text  # the sync API stages its commits
    assert "infinistore_evicted_blocks_total 0" in text
    assert "infinistore_lookup_hits_total 3" in text  # the selftest read three blocks back
    assert "infinistore_dedup_skips_total 0" in text
    assert 'infinistore_op_service_us{op="ALLOCATE",quantile="0.99"}' in text
    lat = stats["op_latency_us"]["ALLOCATE"]
    assert lat["count"] >= 2 and 0 < lat["p50_us"] <= lat["p99_us"] and lat["max_us"] >= lat["mean_us"]
    ckpt = "cli_test.ckpt"
    # checkpoints are bare names inside --checkpoint-dir: a path can never leave it
    for bad in ("/etc/passwd", "../x.ckpt", "..", ".hidden", "a/b"):
        with pytest.raises(urllib.error.HTTPError) as ei:
            _http("POST", base + "/dump?name=" + urllib.parse.quote(bad, safe=""))
        assert ei.value.code == 400
    with pytest.raises(urllib.error.HTTPError) as ei:
        _http("POST", base + "/load?name=absent.ckpt")
    assert ei.value.code == 404
    assert _http("POST", base + f"/dump?name={ckpt}")["num"] == 11
    assert os.path.isfile(os.path.join(CKPT_DIR, ckpt))
    assert not [f for f in os.listdir(CKPT_DIR) if ".tmp." in f]  # written then renamed
    assert _http("POST", base + "/purge") == {"status": "ok", "num": 11}
    assert _http("GET", base + "/kvmap_len") == {"len": 0}
    assert not conn.check_exist("cli-0")
    assert _http("POST", base + f"/load?name={ckpt}")["num"] == 11
    assert conn.check_exist("cli-0")
    dst = torch.zeros(1024)
    conn.read_cache(dst, [("cli-5", 0)], 1024)
    conn.sync()
    assert torch.equal(dst, src[5 * 1024:6 * 1024])
    _http("POST", base + "/purge")


def test_cpu_benchmark_against_cli_server(cli_server):
    sport, _ = cli_server
    from infinistore_b200 import benchmark

    args = benchmark.parse_args(["--service-port", str(sport), "--size", "4", "--block-size", "4",
                                 "--iteration", "2", "--rdma", "--cpu", "--steps", "4"])
    r = benchmark.run(args)
    assert r["write_mb_s"] > 0 and r["read_mb_s"] > 0


def test_arg_defaults_match_reference():
    from infinistore_b200 import server

    a = server.parse_args([])
    assert (a.manage_port, a.service_port, a.prealloc_size, a.minimal_allocate_size) == \
        (18080, 22345, 16, 64)
    assert (a.log_level, a.dev_name, a.ib_port, a.link_type, a.num_stream) == \
        ("info", "mlx5_1", 1, "IB", 1)
    assert a.auto_increase is False and a.warmup is False and a.host == "0.0.0.0"
    cfg = server.config_from_args(server.parse_args(["--pool-devices", "0,2", "--auto-increase"]))
    assert list(cfg.pool_devices) == [0, 2] and cfg.auto_increase
    assert cfg.evict is False and abs(cfg.evict_ratio - 0.05) < 1e-9
    cfg = server.config_from_args(server.parse_args(["--evict", "--evict-ratio", "0.2"]))
    assert cfg.evict is True and abs(cfg.evict_ratio - 0.2) < 1e-9
    cfg.manage_port, cfg.service_port = 1, 2
    cfg.verify()
    cfg.evict_ratio = 0.0
    import pytest
    with pytest.raises(Exception):
        cfg.verify()


def test_checkpoint_policy_token_and_disabled():
    from infinistore_b200.server import CheckpointPolicy

    off = CheckpointPolicy("", "")
    with pytest.raises(PermissionError):
        off.resolve("x.ckpt")  # no --checkpoint-dir: endpoints disabled
    assert off.authorize("127.0.0.1", None) is None
    assert off.authorize("10.1.2.3", None) is not None  # remote callers need a token
    tok = CheckpointPolicy("/tmp/istore_policy_dir", "s3cret")
    assert tok.authorize("10.1.2.3", "s3cret") is None
    assert tok.authorize("127.0.0.1", None) is not None
    assert tok.authorize("127.0.0.1", "wrong") is not None
    assert tok.resolve("a.ckpt") == os.path.realpath("/tmp/istore_policy_dir/a.ckpt")
    for bad in ("../a", "/abs", "a/b", ".a", "", "a..b"):
        with pytest.raises(ValueError):
            tok.resolve(bad)

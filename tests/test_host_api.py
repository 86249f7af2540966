"""Host-side mirror of the Rust API (clann_b200/api.py): Config semantics, error mapping, recall helper."""
import numpy as np
import pytest

import clann_b200 as cb


def test_default_config():  # src/core/config.rs:38-47 (the reference's own test asserts num_tables == 1 and fails)
    c = cb.Config()
    assert (c.num_tables, c.num_clusters_factor, c.k, c.delta, c.dataset_name) == (10, 1.0, 10, 0.9, "")
    assert c.metrics_output is cb.MetricsOutput.NONE


def test_new_config_and_json():  # config.rs:85-126
    c = cb.Config.new(2048, 10.0, 100, 0.95, "test_dataset", cb.MetricsOutput.NONE)
    assert (c.num_tables, c.num_clusters_factor, c.k, c.delta, c.dataset_name) == (2048, 10.0, 100, 0.95, "test_dataset")
    j = c.to_json_dict()
    assert j["num_tables"] == 2048 and j["num_clusters_factor"] == 10.0 and j["metrics_output"] == "None"


def test_empty_dataset_is_data_error():  # index.rs:72-74
    with pytest.raises(cb.DataError, match="empty dataset"):
        cb.init_with_config(np.zeros((0, 8), np.float32), cb.Config(4, 1.0, 3, 0.9))


def test_bad_config_is_config_error():
    with pytest.raises(cb.ConfigError):
        cb.init_with_config(np.zeros((5, 8), np.float32), cb.Config(0, 1.0, 3, 0.9))


def test_error_messages_follow_errors_rs():  # src/core/errors.rs:6-39
    assert str(cb.DataError("empty dataset")) == "Data Error: empty dataset"
    assert str(cb.PuffinnCreationError("x")) == "PUFFINN Creation Error: x"
    assert str(cb.IndexNotFound()) == "Index Not Found Error"


def test_recall_values():  # src/utils/mod.rs:59-95
    truth = np.array([[0.1, 0.2, 0.3, 0.9], [0.0, 0.5, 0.6, 0.7]], np.float32)
    run = [[0.1, 0.2, 0.3005], [0.0, 0.5, 0.8]]
    mean, std, per = cb.get_recall_values(truth, run, 3)
    assert per == [3.0, 2.0] and abs(mean - 5 / 6) < 1e-6


def test_unit_vector_helper():
    v = cb.generate_random_unit_vectors(50, 16, seed=1)
    assert v.shape == (50, 16) and np.allclose(np.linalg.norm(v, axis=1), 1.0, atol=1e-5) and (v >= 0).all()


def test_run_metrics_rows(tmp_path):
    """RunMetrics (utils/metrics/mod.rs:14-35,254-262): queries/s = queries / total time, recall by get_recall_values."""
    import clann_b200 as cb
    from clann_b200 import api
    cfg = cb.Config(84, 0.4, 2, 0.9, "unit")
    ctr = dict(distance_computations=[5, 7, 9], candidates=[50, 70, 90], clusters_visited=[1, 2, 1])
    gt = np.array([[0.1, 0.2, 0.3], [0.1, 0.2, 0.3], [0.5, 0.6, 0.7]], np.float32)
    run = [[0.1, 0.2], [0.1, 0.9], [0.9, 0.95]]
    m = api.RunMetrics(cfg, 1000, 0.5, ctr, run, gt)
    assert m.queries_per_second == 6.0
    assert abs(m.recall_mean - 0.5) < 1e-6 and m.recalls == [2.0, 1.0, 0.0]
    rows = m.query_rows()
    assert rows[1]["distance_computations"] == 7 and rows[1]["n_candidates"] == 70 and rows[1]["recall"] == 0.5
    assert m.run_row()["num_tables"] == 84 and m.run_row()["dataset_len"] == 1000
    p = tmp_path / "m.csv"
    m.write_csv(str(p))
    lines = p.read_text().splitlines()
    assert lines[1].startswith("query_idx,") and len(lines) == 5 and lines[3].split(",")[2] == "7"
    import json
    assert len(json.loads(m.to_json())["queries"]) == 3


def test_record_container_round_trip(tmp_path):
    """The flat record container of serialize / CPUFFINN_save_index: append-only, the last record of a name wins, foreign
    files and truncated records are SerializeError; a missing file is ConfigError (index.rs:108-113)."""
    import clann_b200 as cb
    from clann_b200 import api
    p = tmp_path / "f.clb2"
    with open(p, "wb") as f:
        api._write_record(f, "config", b'{"a": 1}')
        api._write_record(f, "index_3", bytes(range(200)))
        api._write_record(f, "config", b'{"a": 2}')
    rec = api._read_records(str(p))
    assert rec["config"] == b'{"a": 2}' and rec["index_3"] == bytes(range(200)) and set(rec) == {"config", "index_3"}
    raw = p.read_bytes()
    (tmp_path / "t.clb2").write_bytes(raw[:-10])
    with pytest.raises(cb.api.SerializeError):
        api._read_records(str(tmp_path / "t.clb2"))
    (tmp_path / "x.clb2").write_bytes(b"not a record file at all")
    with pytest.raises(cb.api.SerializeError):
        api._read_records(str(tmp_path / "x.clb2"))
    with pytest.raises(cb.api.ConfigError):
        api.init_from_file(np.zeros((4, 4), np.float32), str(tmp_path / "missing.clb2"))


def test_bench_reference_arm_contract():
    """bench.py --impl reference under torchrun: rank 0 alone runs, the other ranks exit 0 without work and without output;
    the data generators of the two arms are the same functions (same seeds, same shapes)."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT="29999")
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "2",
                        "--warmup", "1"], env=env, capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and r.stdout.strip() == ""
    sys.path.insert(0, root)
    import argparse
    import bench
    w = bench.workload(argparse.Namespace(workload="glove100", small=True, rows=0, delta=0.0))
    d1, q1, s1 = bench.make_data(w, "planted")
    d2, q2, s2 = bench.make_data(w, "planted")
    assert d1.shape == (w["n"], w["d"]) and q1.shape == (w["nq"], w["d"]) and d1.dtype == np.float32
    assert np.array_equal(d1, d2) and np.array_equal(q1, q2) and np.array_equal(s1, s2)
    assert np.allclose(np.linalg.norm(d1, axis=1), 1.0, atol=1e-5)

"""Host-side pieces of bench.py that do not need a GPU: the MEASURED_PEAKS.json reader (the file is written by the
driver, its schema is not ours) and the workload table against BASELINE.json's configs."""
import json
import os

import bench


def test_workloads_match_the_baseline_configs():
    with open(os.path.join(bench.ROOT, "BASELINE.json")) as f:
        configs = json.load(f)["configs"]
    assert len(configs) == 5
    w = bench.WORKLOADS
    assert w["cfg1"][:2] == ("float32", (1 << 24,))            # 1D fp32 16 Mi
    assert w["cfg2"][:2] == ("float32", (512, 512, 512))       # 3D fp32 512^3 (the bench line's workload)
    assert w["cfg3"][:2] == ("float64", (8192, 8192))          # 2D fp64 8192^2
    assert w["cfg4"][:2] == ("float64", (128, 1024, 1024))     # 3D fp64 1024^3 / 8 ranks
    assert w["cfg5"][:2] == ("float32", (1 << 28,))            # 1D fp32 2 Gi / 8 ranks


def test_measured_peak_reader_accepts_unknown_schemas(tmp_path, monkeypatch):
    monkeypatch.setattr(bench, "ROOT", str(tmp_path))
    assert bench.measured_hbm_peak()[0] == 6650.0              # file absent: the profiling guide's fallback
    cases = [({"hbm_gbs": 6553.6, "bf16_tflops": 1500}, 6553.6),
             ({"hbm": {"burst_gbs": 6700, "sustained_gbs": 6400}}, 6700.0),
             ({"peaks": {"hbm_copy_tbs": 6.55}}, 6550.0),
             ({"foo": 1}, 6650.0)]
    for doc, want in cases:
        (tmp_path / "MEASURED_PEAKS.json").write_text(json.dumps(doc))
        value, source = bench.measured_hbm_peak()
        assert abs(value - want) < 1e-6, (doc, value, source)
    (tmp_path / "MEASURED_PEAKS.json").write_text("not json")
    assert bench.measured_hbm_peak()[0] == 6650.0

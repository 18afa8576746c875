"""CPU: bench.py's contract pieces that do not need a GPU -- stdout carries nothing but the JSON line (libraries that
print to stdout are diverted), and without a GPU the product arm fails loudly instead of printing a number."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_stdout_is_reserved_for_the_json_line():
    code = (
        "import os, sys; sys.path.insert(0, %r); import bench\n"
        "bench.claim_stdout()\n"
        "os.write(1, b'NCCL version banner\\n')\n"      # what a library writing to fd 1 does
        "print('python print to stdout')\n"
        "bench.emit({'metric': 'Msamples/s', 'value': 1.0})\n" % ROOT
    )
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, check=True)
    lines = out.stdout.splitlines()
    assert len(lines) == 1 and json.loads(lines[0]) == {"metric": "Msamples/s", "value": 1.0}
    assert "NCCL version banner" in out.stderr and "python print to stdout" in out.stderr


def test_product_arm_fails_loudly_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("a GPU is present")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "0"], capture_output=True, text=True)
    assert out.returncode != 0
    assert out.stdout.strip() == ""


def test_workload_table_and_roofline_inputs():
    sys.path.insert(0, ROOT)
    import bench
    import ky_b200 as ky
    wl = bench.workloads(ky)
    assert sorted(wl) == ["C1", "C2", "C3", "C4", "C5"]
    assert [len(wl[c]["panels"]) for c in ("C1", "C2", "C3", "C4", "C5")] == [1, 3, 1, 12, 1]
    assert sum(p[2] * p[3] for p in wl["C4"]["panels"]) == 1920 * 1080
    # 712 flop per ray in the Cornell default scene, 488 in the Veach scene, 144 in the smallpt scene
    assert bench.flop_per_ray(ky.Scene(ky.SCENE_CORNELL, 8, 8, ky.CB_DEFAULT)) == 712
    assert bench.flop_per_ray(ky.Scene(ky.SCENE_VEACH, 8, 8)) == 488
    assert bench.flop_per_ray(ky.Scene(ky.SCENE_SMALLPT, 8, 8)) == 144


def test_ncu_summary_must_describe_this_build(tmp_path, monkeypatch):
    """The ncu figures the line quotes come from a committed capture; one that names kernels the library does not contain is refused."""
    sys.path.insert(0, ROOT)
    import bench
    prof = tmp_path / "profiles"
    prof.mkdir()
    lib = tmp_path / "ky_b200" / "lib"
    lib.mkdir(parents=True)
    (lib / "libkyd.so").write_bytes(open(os.path.join(ROOT, "ky_b200", "lib", "libkyd.so"), "rb").read())
    csrc = tmp_path / "ky_b200" / "csrc"
    csrc.mkdir()
    (csrc / "a.cu").write_text("x")
    monkeypatch.setattr(bench, "ROOT", str(tmp_path))
    assert bench.ncu_summary() is None
    doc = {"source_hash": bench.source_hash(), "kernel_symbols": ["k_intersect", "k_shade"], "kernels": {"k_shade<lambert>": {"dram_bytes": 1.0}}}
    (prof / "r02_ncu_summary.json").write_text(json.dumps(doc))
    got = bench.ncu_summary()
    assert got["same_sources"] is True and got["kernels"]["k_shade<lambert>"]["dram_bytes"] == 1.0
    (csrc / "a.cu").write_text("y")
    assert bench.ncu_summary()["same_sources"] is False
    doc["kernel_symbols"].append("k_no_such_kernel")
    (prof / "r02_ncu_summary.json").write_text(json.dumps(doc))
    import pytest
    with pytest.raises(SystemExit):
        bench.ncu_summary()

"""The measurement contract, checked without a GPU: bench.py keeps SURVEY.md section 8(d)'s byte model (the traffic of
the reference's structure) beside the model of what THIS design moves, which is the one every printed GB/s uses; the
roofline denominator is the driver-measured peak, and the bench line recorded on the B200 (profiles/r01_v6_bench.json)
carries every key the driver reads."""
import json
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def bench():
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_module", os.path.join(ROOT, "bench.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_survey_bytes_follow_survey_8d(bench):
    # SURVEY 8(d) worked example: headline config with V = 0.8 P = 2.4 M, N = 6 V = 14.4 M
    P, V, N, W, H, M = 3_000_000, 2_400_000, 14_400_000, 1600, 1008, 16
    G = (W // 16) * (H // 16)
    a = bench.survey_bytes(P, V, N, G, W, H, M)
    assert a["preprocess"] == P * (119 + 12 * M) == P * 311
    assert a["scan"] == 8 * P and a["duplicate"] == 20 * P + 12 * N
    assert a["tile_sort"] == N * (8 + 6 * 24)                       # 1 histogram read + 6 passes x (12 read + 12 write)
    assert a["tile_ranges"] == 8 * N + 8 * G
    assert a["blend_forward"] == 44 * N + 24 * W * H
    assert a["blend_backward"] == 44 * N + 20 * W * H + 36 * N
    assert a["geom_backward"] == P * 550
    fwd = sum(a[k] for k in ("preprocess", "scan", "duplicate", "tile_sort", "tile_ranges", "blend_forward"))
    bwd = a["blend_backward"] + a["geom_backward"]
    assert abs(fwd / 1e9 - 4.2) < 0.15 and abs(bwd / 1e9 - 2.8) < 0.15          # "~4.2 GB + ~2.8 GB = ~7 GB per view"
    for M_, b in ((1, 190), (4, 260)):
        assert bench.survey_bytes(10, 10, 10, 1, 16, 16, M_)["geom_backward"] == 10 * b
    assert set(a) <= set(bench.FWD_STAGES) | set(bench.BWD_STAGES)         # every modelled stage is a profiled stage


def test_design_bytes_describe_this_design(bench):
    """The model behind every printed GB/s (round-1 verdict: SURVEY's six-pass 64-bit sort was charged to a two-pass
    32-bit one and the batched K8+K9 was charged per view -- fractions above 1.0).  Headline numbers as measured."""
    P, V, N, W, H, M = 3_000_000, 1_790_869, 8_209_874, 1600, 1008, 16
    G = (W // 16) * (H // 16)
    one, four = bench.algorithmic_bytes(P, V, N, G, W, H, M, 1), bench.algorithmic_bytes(P, V, N, G, W, H, M, 4)
    assert set(one) <= set(bench.FWD_STAGES) | set(bench.BWD_STAGES)
    assert one["tile_sort"] == 2 * 16 * N                                 # 13-bit tile ids: two 8-bit passes of 16 B pairs
    assert one["depth_sort"] == 8 * P + 56 * V                            # histogram + compacting pass + three passes over V
    assert one["duplicate"] == 16 * V + 8 * N and one["accum_clear"] == 48 * P
    assert one["blend_forward"] == 44 * N + 24 * W * H and one["blend_backward"] == 80 * N + 20 * W * H
    # views of one batched launch share the per-Gaussian reads (and, in K8+K9, the gradient writes)
    assert four["preprocess"] < one["preprocess"] and four["geom_backward"] < one["geom_backward"]
    assert one["preprocess"] - four["preprocess"] == pytest.approx(0.75 * (44 * P + 12 * M * V))
    # the design moves far less than the reference's structure on the stages it restructured
    sv = bench.survey_bytes(P, V, N, G, W, H, M)
    assert one["tile_sort"] + one["depth_sort"] + one["duplicate"] < 0.5 * (sv["tile_sort"] + sv["duplicate"] + sv["scan"])
    # a G with more than 16 bits of tile id needs three passes
    assert bench.algorithmic_bytes(10, 10, 100, 70000, 4000, 4500, 1)["tile_sort"] == 3 * 16 * 100


def test_peak_is_the_measured_one(bench):
    peak, kind = bench.peaks()
    mp = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(mp):
        assert kind == "measured" and peak == float(json.load(open(mp))["hbm_gbs"])
    else:
        assert kind == "fallback" and 6000 < peak < 8000


def test_recorded_bench_line_has_the_contract_keys():
    line = json.loads(open(os.path.join(ROOT, "profiles", "r01_v6_bench.json")).read().strip().splitlines()[-1])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "clocks", "roofline", "cpu_baseline"):
        assert k in line, k
    assert line["unit"] == "views/s" and line["higher_is_better"] is True and line["scaling"] == "weak"
    assert line["vs_baseline"] is None and line["data"] == "synthetic" and line["dtype"] == "f32"
    assert line["warmup"] >= 3 and line["steps"] >= 20 and line["gpu_launches"] > 0
    assert "workload" in line["config"] and "model" not in line["config"]
    assert {"P", "V", "N", "G"} <= set(line["config"])                           # so the roofline can be recomputed
    e = line["e2e"]
    assert e["unit"] == line["unit"] and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0
    assert 0 < e["value"] < line["value"] * 1.02                                 # end to end is not a copy of `value`
    r = line["roofline"]
    assert r["bound"] in ("hbm", "tensor") and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert r["traffic"] is None or r["traffic"] > 0
    c = line["cpu_baseline"]
    assert c["kind"] in ("reference", "port") and c["cores"] >= 1 and c["unit"] == line["unit"] and c["sample"]
    k = line["clocks"]
    assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(k)
    assert not set(k["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    # value = views of the whole job / device time
    assert abs(line["value"] - line["config"]["views_per_step"] / (line["ms_per_step"] / 1000.0)) < 1e-6 * line["value"]


def test_view_roofline_helper():
    """the whole-view roofline object of the bench line: design bytes without the L2-served list gathers, and SURVEY 8(d)'s
    reference-structure B_view beside it (pure arithmetic: checked here, the bench itself needs a GPU)"""
    import bench
    P, V, N, G, W, H, M = 3_000_000, 1_790_869, 5_733_709, 6300, 1600, 1008, 16
    alg = bench.algorithmic_bytes(P, V, N, G, W, H, M, 4, clear_in_k1=True)
    sv = bench.survey_bytes(P, V, N, G, W, H, M)
    r = bench.view_roofline(alg, N, W, H, 1.386, 6537.6, survey=sv)
    non_blend = sum(v for k, v in alg.items() if not k.startswith("blend"))
    assert r["alg_bytes"] == int(non_blend + 44 * W * H) and r["blend_list_gather_upper_bound_bytes"] == 124 * N
    assert 0.1 < r["frac"] < 0.2 and abs(r["gbps"] - r["alg_bytes"] / 1.386e-3 / 1e9) < 1e-6
    s8 = r["survey_8d"]
    assert s8["bytes"] == int(sum(sv.values())) and 0.3 < s8["equivalent_frac_of_hbm_peak"] < 0.8
    assert "survey_8d" not in bench.view_roofline(alg, N, W, H, 1.386, 6537.6)

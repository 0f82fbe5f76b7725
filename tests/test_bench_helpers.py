"""CPU checks of bench.py's host-side bookkeeping (no GPU, no timing)."""
import importlib.util
import os

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def _bench():
    spec = importlib.util.spec_from_file_location("bench", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_workload_geometry_matches_the_shape_logic_and_survey_bytes():
    import biahub_b200 as b2

    bench = _bench()
    for name, w in bench.WORKLOADS.items():
        out_shape, bytes_unit, out_vox = bench.unit_geometry(w)
        if w["kind"] in ("deskew", "chain"):
            want, _ = b2.get_deskewed_data_shape(w["shape"], w["ls_angle_deg"], w["px_to_scan_ratio"],
                                                 w["keep_overhang"], w["average_n_slices"])
            assert tuple(want) == out_shape, name
        else:
            assert out_shape == tuple(w["shape"]), name
        assert out_vox == int(np.prod(out_shape))
    # SURVEY.md §8(d) figures
    assert bench.unit_geometry(bench.WORKLOADS["deskew_c2"])[1] == 2_468_249_600
    assert bench.unit_geometry(bench.WORKLOADS["register_c3"])[1] == 4_026_531_840
    assert bench.unit_geometry(bench.WORKLOADS["stabilize_c4"])[1] == 2_147_483_648
    assert bench.unit_geometry(bench.WORKLOADS["deskew_c1"])[0] == (256, 512, 442)


def test_plate_partition_covers_every_unit_once():
    from biahub_b200.sharding import enumerate_units, units_for_rank

    bench = _bench()
    w = bench.WORKLOADS["plate_c5"]
    units = enumerate_units(w["positions"], range(w["timepoints"]), [0])
    assert len(units) == 256
    for world in (1, 2, 4, 8):
        shards = [units_for_rank(units, r, world) for r in range(world)]
        assert sorted(u for s in shards for u in s) == sorted(units)
        assert max(map(len, shards)) - min(map(len, shards)) <= 1


def test_reference_arm_config_equals_gpu_arm_config_keys():
    bench = _bench()
    w = bench.WORKLOADS["deskew_c2"]
    out_shape, bytes_unit, _ = bench.unit_geometry(w)
    # the keys measure() puts into `config`
    gpu_keys = {"workload", "source_dtype", "units_per_step_per_gpu", "out_shape", "sharding", "l2"}
    import inspect

    src = inspect.getsource(bench.run_reference)
    for k in gpu_keys:
        assert f'"{k}"' in src, k

"""CPU checks of the measurement contract and of the binding tables: both bench arms describe
the workload identically, and every C entry point has a Python binding."""
import json
import os
import re

import amgcl_b200 as ab

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PROF = os.path.join(ROOT, "profiles")


def load(name):
    with open(os.path.join(PROF, name)) as f:
        return json.loads(f.read())


def test_both_bench_arms_describe_the_workload_with_the_same_keys():
    """The driver compares the `config` of `bench.py` and `bench.py --impl reference`: both are
    built by bench.config_block, so their key sets and the workload string are identical."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)

    class Args:
        n, relax, krylov, precision = 256, "damped_jacobi", "cg", "f64"
    ours = bench.config_block(Args, 10, 20, 1.0, 2.0, backend="amgcl::backend::b200<double>",
                              parallelism="single GPU")
    ref = bench.config_block(Args, 10, 20, 3.0, 4.0, backend="amgcl::backend::builtin<double> (OpenMP)")
    assert set(ours) == set(ref)
    assert ours["workload"] == ref["workload"] == "poisson3d_256^3_fp64_sa_damped_jacobi_cg"
    assert "model" not in ours and ours["l2"].startswith("inputs_exceed_l2")


def test_every_c_entry_point_has_a_python_binding():
    """include/amgcl_b200.h is the boundary; amgcl_b200/__init__.py binds all of it (the tests
    call the product through these bindings only)."""
    header = open(os.path.join(ROOT, "include", "amgcl_b200.h")).read()
    declared = set(re.findall(r"\b(b200_[a-z0-9_]+)\s*\(", header))
    declared -= {"b200_ctx_s", "b200_profile_entry"}
    src = open(os.path.join(ROOT, "amgcl_b200", "__init__.py")).read()
    bound = set(re.findall(r"\b(b200_[a-z0-9_]+)\b", src))
    missing = sorted(declared - bound)
    assert not missing, missing
    L = ab.lib()
    for name in declared:
        assert hasattr(L, name), name

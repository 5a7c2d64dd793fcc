"""CPU-side checks of the drop-in boundary: the C-ABI library loads without a GPU and exports every entry
point include/jampr_b200.h declares; struct layouts of the ctypes binding match the header; the host-side
mirrors refuse to run without CUDA (no fallback)."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "jampr_b200.h")


def _declared():
    src = open(HEADER).read()
    return re.findall(r"JAMPR_API\s+[\w\s\*]+?\b(jampr_\w+)\s*\(", src)


def test_library_exports_every_declared_symbol():
    from jampr_plus_b200 import _cabi
    lib = _cabi.load_library()
    names = _declared()
    assert len(names) >= 13
    for n in names:
        assert hasattr(lib, n), n
    assert set(names) == set(_cabi.SYMBOLS)
    assert lib.jampr_abi_version() == _cabi.ABI_VERSION


def test_struct_fields_match_header():
    from jampr_plus_b200 import _cabi
    src = open(HEADER).read()

    def fields(struct_name):
        end = src.index("} " + struct_name + ";")
        body = src[src.rindex("typedef struct {", 0, end) + len("typedef struct {"): end]
        body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
        return [re.search(r"(\w+)\s*;", line).group(1) for line in body.split("\n") if ";" in line]

    assert fields("jampr_dims_t") == [f[0] for f in _cabi.Dims._fields_]
    assert fields("jampr_instances_t") == [f[0] for f in _cabi.Instances._fields_]
    assert fields("jampr_state_t") == [f[0] for f in _cabi.State._fields_]
    assert fields("jampr_weights_t") == [f[0] for f in _cabi.Weights._fields_]
    assert ctypes.sizeof(_cabi.Dims) == 32


def test_weight_blob_layout_and_packing_roundtrip():
    from jampr_plus_b200 import _cabi
    from jampr_plus_b200.packing import pack_f32, validate_state_dict
    from jampr_plus_b200.weights import random_state_dict, state_dict_schema
    off = _cabi.weight_offsets()
    lib = _cabi.load_library()
    assert off["END"] == lib.jampr_weight_blob_floats()
    sd = random_state_dict(0)
    blob = pack_f32(sd)
    assert blob.numel() == off["END"]
    # spot checks of the layout the kernels assume (csrc/common.cuh)
    w = sd["node_encoder.layers.1.conv.lin_rel.weight"]
    blk = off["NE_BLK"] + off["BLK_FLOATS"]
    assert torch.equal(blob[blk: blk + 128 * 128].view(128, 128), w.t())
    sp = sd["decoder.set_proj.weight"]
    assert torch.equal(blob[off["GV_WT"]: off["GV_WT"] + 65536].view(256, 256), sp[256:512].t())
    # older PyG key aliases are accepted, unknown keys are not
    alias = {k.replace("lin_rel", "lin_l").replace("lin_root", "lin_r"): v for k, v in sd.items()}
    assert set(validate_state_dict(alias)) == set(state_dict_schema())
    bad = dict(sd)
    bad["decoder.extra"] = torch.zeros(1)
    with pytest.raises(KeyError):
        validate_state_dict(bad)


def test_no_cpu_fallback():
    from jampr_plus_b200 import Policy, RPEnv
    with pytest.raises(RuntimeError):
        RPEnv(device="cpu")
    pol = Policy(RPEnv.OBSERVATION_SPACE, device="cpu")
    from jampr_plus_b200.routing.formats import RPObs
    obs = RPObs(1, None, None, None, None, None, None, None, None, None)
    with pytest.raises(RuntimeError):
        pol(obs)


def test_policy_state_dict_schema_matches_reference():
    import json
    from jampr_plus_b200 import Policy, RPEnv
    sch = json.load(open(os.path.join(ROOT, "tests", "golden", "state_dict_schema.json")))
    pol = Policy(RPEnv.OBSERVATION_SPACE, device="cpu")
    got = {k: list(v.shape) for k, v in pol.state_dict().items()}
    assert list(got.keys()) == list(sch.keys())
    assert got == sch


def test_policy_refuses_cpu_and_checks_schema():
    """No fallback path: constructing the mirrors needs the library, running them needs CUDA; load_state_dict accepts
    the reference schema (old PyG aliases included) and nothing else."""
    from jampr_plus_b200 import Policy, RPEnv
    from jampr_plus_b200.weights import random_state_dict
    with pytest.raises(RuntimeError):
        RPEnv(device="cpu")
    pol = Policy(RPEnv.OBSERVATION_SPACE, device="cpu")      # parameter container only; forward needs a CUDA env
    sd = random_state_dict(2)
    alias = {k.replace(".conv.lin_rel.", ".conv.lin_l.").replace(".conv.lin_root.", ".conv.lin_r."): v for k, v in sd.items()}
    pol.load_state_dict(alias)
    assert torch.equal(pol.state_dict()["node_encoder.layers.0.conv.lin_rel.weight"], sd["node_encoder.layers.0.conv.lin_rel.weight"])
    bad = dict(sd)
    bad["decoder.extra.weight"] = torch.zeros(1)
    with pytest.raises(RuntimeError):
        pol.load_state_dict(bad)
    with pytest.raises(NotImplementedError):
        Policy(RPEnv.OBSERVATION_SPACE, device="cpu", embedding_dim=128)


def test_header_is_plain_c_and_cxx(tmp_path):
    """include/jampr_b200.h is the drop-in boundary: it must compile as C99 and as C++ without CUDA or torch headers,
    and a C translation unit that takes the address of every entry point must link against the library."""
    import re
    import shutil
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    hdr = os.path.join(root, "include", "jampr_b200.h")
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only", "-x", "c", hdr],
                   check=True)
    subprocess.run(["g++", "-std=c++17", "-Wall", "-Werror", "-fsyntax-only", "-x", "c++", hdr], check=True)
    names = sorted(set(re.findall(r"\b(jampr_[a-z0-9_]+)\s*\(", open(hdr).read())))
    assert len(names) >= 15
    src = tmp_path / "link.c"
    src.write_text('#include "jampr_b200.h"\n#include <stdio.h>\nint main(void) {\n  void* p[] = {'
                   + ", ".join(f"(void*){n}" for n in names)
                   + '};\n  printf("%d %d\\n", (int)(sizeof(p) / sizeof(p[0])), jampr_abi_version());\n  return 0;\n}\n')
    lib_dir = os.path.join(root, "jampr_plus_b200", "lib")
    exe = tmp_path / "link"
    subprocess.run(["gcc", "-std=c99", "-I", os.path.join(root, "include"), str(src), "-o", str(exe),
                    "-L", lib_dir, "-ljampr_b200", f"-Wl,-rpath,{lib_dir}", "-Wno-pedantic"], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()
    from jampr_plus_b200 import _cabi
    assert int(out[0]) == len(names) and int(out[1]) == _cabi.ABI_VERSION

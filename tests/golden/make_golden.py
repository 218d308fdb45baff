"""Generates tests/golden/*.npz with the UNMODIFIED reference (oracle/_ref/libcmref.so).

Run in the build container, where /root/reference exists:
    make -C oracle ref && python tests/golden/make_golden.py
Each fixture holds the BSP bytes, the query inputs and the reference's outputs (56-byte trace_t
records as raw bytes, or contents words), so the oracle and the CUDA engine can be checked on
machines that have neither the reference tree nor the reference library."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import cases  # noqa: E402
from helpers import cmref, get_map  # noqa: E402

N = 1500


def dump(name, kind, case_list):
    m = get_map(kind)
    r = cmref.RefCM().load(m.data)
    blob = {"bsp": np.frombuffer(m.data, dtype=np.uint8), "num_cases": np.array(len(case_list))}
    for i, c in enumerate(case_list):
        out = cases.run_case(r, c)
        blob[f"c{i}_name"] = np.array(c["name"])
        blob[f"c{i}_kind"] = np.array(c["kind"])
        for k, v in c.items():
            if k in ("name", "kind", "map") or v is None:
                continue
            blob[f"c{i}_in_{k}"] = np.asarray(v)
        blob[f"c{i}_out"] = out.view(np.uint8).reshape(len(out), -1) if out.dtype.names else out
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **blob)
    print(path, os.path.getsize(path) // 1024, "KiB", len(case_list), "cases")


if __name__ == "__main__":
    dump("tiny_world", "tiny", cases.world_cases("tiny", n=N))
    dump("room_world", "room", cases.world_cases("room", n=N))
    dump("models_world", "arena_models", cases.world_cases("arena_models", n=N)[:4])
    dump("models_entities", "arena_models", cases.entity_cases("arena_models", n=N))
    dump("patches_world", "patches_small", cases.world_cases("patches_small", n=N)[:6])
    dump("patches_entities", "patches_small", cases.entity_cases("patches_small", n=N)[:3])

"""CPU, static: every kernel that libpcbridge launches with the programmatic-dependent-launch attribute (`launch_pdl`,
csrc/pcb_common.cuh) must call `pdl_wait()` -- griddepcontrol.wait -- before it touches global memory, and must trigger
its dependents only AFTER that wait (wait-then-trigger keeps the overlap one kernel deep).  A kernel launched with the
attribute but without the wait would start on its predecessor's unfinished output."""
import os
import re

CSRC = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "pointcloud_bridge_b200", "csrc")


def _kernel_body(src: str, name: str) -> str:
    m = re.search(r"\n" + re.escape(name) + r"\((?:[^{;]|\n)*?\)\n\{\n", src)
    assert m, f"definition of {name} not found"
    depth, i = 1, m.end()
    while depth and i < len(src):
        depth += {"{": 1, "}": -1}.get(src[i], 0)
        i += 1
    return src[m.end():i]


def test_every_pdl_launched_kernel_waits_before_it_triggers():
    launched = {}
    for f in sorted(os.listdir(CSRC)):
        if f.endswith(".cu"):
            src = open(os.path.join(CSRC, f)).read()
            for name in set(re.findall(r"launch_pdl\(\s*([A-Za-z_0-9]+)", src)):
                launched[name] = (f, src)
    assert len(launched) >= 14, sorted(launched)
    for name, (f, src) in launched.items():
        body = _kernel_body(src, name)
        w, t = body.find("pdl_wait();"), body.find("pdl_trigger();")
        assert w >= 0, f"{f}: {name} is launched with the PDL attribute but never waits"
        assert t < 0 or w < t, f"{f}: {name} triggers its dependents before its own wait"
        # nothing before the wait may read global memory: only CTA-local setup (shared memory, TMEM, barriers, the
        # kernel's own parameters) -- a plain heuristic: no __ldg / ld.global / pointer dereference of a parameter array
        head = body[:w]
        assert "__ldg(" not in head and "ld.global" not in head, f"{f}: {name} loads from global memory before pdl_wait()"

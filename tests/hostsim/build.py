"""TEST INFRASTRUCTURE: compile the CUDA-core kernels of a ccdm_b200/csrc/*.cu file for the HOST (g++, C++20) through
tests/hostsim/cuda_host_shim.h, so that the kernel source itself -- not a restatement of it -- runs on the CPU.

The .cu text is used verbatim except for (1) its two project includes, which the shim replaces, and (2) the
``kernel<<<grid, block, 0, stream>>>(args);`` launch syntax, rewritten to ``LAUNCH(kernel, (grid), (block), args);``, and (3) ``extern __shared__ float name[];``, which becomes a pointer to the shim's
dynamic-shared-memory buffer.
"""
import hashlib
import os
import re
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
LAUNCH_RE = re.compile(r"([A-Za-z_]\w*(?:<[\w, ]+>)?)<<<(.*?),\s*([^,]+?),\s*([^,]+?),\s*([^,]*?)>>>\((.*?)\);", re.S)
DYN_SMEM_RE = re.compile(r"extern\s+__shared__\s+float\s+(\w+)\[\];")


SYNC_RE = re.compile(r"__syncthreads|__shfl|__syncwarp|block_sum|warp_sum|seg_sum")


def _launch_sub(src: str):
    """Rewrites the launches of ``src``.  A kernel whose body uses no barrier / shuffle (directly or through the project's
    reduction helpers) does not need concurrent threads: its CUDA threads run one after the other in the calling thread
    (LAUNCH_SEQ), which is much faster than a fiber each."""
    def repl(m):
        base = m.group(1).split("<")[0]
        body = _definition(src, r"__global__ void (?:__launch_bounds__\([^)]+\) )?" + base + r"\(")
        macro = "LAUNCH" if SYNC_RE.search(body) else "LAUNCH_SEQ"
        return f"{macro}(({m.group(1)}), ({m.group(2)}), ({m.group(3)}), {m.group(6)});"
    return LAUNCH_RE.subn(repl, src)


def host_source(cu_path: str) -> str:
    src = open(cu_path).read()
    src = src.replace('#include "common.cuh"', "").replace('#include "ptx.cuh"', "")
    src = DYN_SMEM_RE.sub(r"float* \1 = g_dyn_smem;", src)          # dynamic shared memory: one 256 KB host buffer
    src, n = _launch_sub(src)
    assert n > 0 and "<<<" not in src, "unconverted kernel launch"
    common = open(os.path.join(ROOT, "ccdm_b200", "csrc", "common.cuh")).read()
    i = common.index("inline void row_lane_plan")
    plan_fn = common[i:common.index("\n}\n", i) + 3]                     # the host helper, verbatim from common.cuh
    return '#include "cuda_host_shim.h"\nnamespace ccdm {\n' + plan_fn + "}\n" + src


def _definition(src: str, header_re: str) -> str:
    """Text of the function whose header matches ``header_re`` (from the start of that line to the matching brace)."""
    m = re.search(header_re, src)
    assert m, header_re
    start = src.rfind("\n", 0, m.start()) + 1
    i = src.index("{", m.end())
    depth = 0
    while True:
        depth += {"{": 1, "}": -1}.get(src[i], 0)
        i += 1
        if depth == 0:
            return src[start:i] + "\n"


def extract_source(cu_path: str, kernels, entries) -> str:
    """A .cu file that also holds tcgen05 / TMA code cannot be compiled for the host as a whole: take only the named
    CUDA-core kernels and C entry points (verbatim) out of it."""
    src = open(cu_path).read()
    body = "namespace ccdm {\n" + "".join(_definition(src, r"__global__ void (?:__launch_bounds__\([^)]+\) )?" + k + r"\(")
                                          for k in kernels) + "}\nusing namespace ccdm;\n"
    body += "".join(_definition(src, r'extern "C" int ' + e + r"\(") for e in entries)
    body = DYN_SMEM_RE.sub(r"float* \1 = g_dyn_smem;", body)       # dynamic shared memory: one 256 KB host buffer
    body, n = _launch_sub(body)
    assert n > 0 and "<<<" not in body
    return '#include "cuda_host_shim.h"\n' + body


def build_extract(cu_name: str, kernels, entries) -> str:
    cu_path = os.path.join(ROOT, "ccdm_b200", "csrc", cu_name)
    text = extract_source(cu_path, kernels, entries)
    return _compile(text, os.path.splitext(cu_name)[0] + "_part")


def _compile(text: str, stem: str) -> str:
    tag = hashlib.sha1((text + open(os.path.join(HERE, "cuda_host_shim.h")).read()).encode()).hexdigest()[:12]
    out_dir = os.path.join(HERE, "build")
    os.makedirs(out_dir, exist_ok=True)
    so = os.path.join(out_dir, f"{stem}_host_{tag}.so")
    if not os.path.exists(so):
        cpp = so[:-3] + ".cpp"
        with open(cpp, "w") as f:
            f.write(text)
        subprocess.run(["g++", "-std=c++20", "-O1", "-U_FORTIFY_SOURCE", "-D_FORTIFY_SOURCE=0", "-fPIC", "-shared", "-I", HERE,
                        "-I", os.path.join(ROOT, "include"), cpp, "-o", so], check=True)
    return so


def build(cu_name: str) -> str:
    """Returns the path of the host-compiled shared library for ccdm_b200/csrc/<cu_name> (rebuilt when the source changes)."""
    cu_path = os.path.join(ROOT, "ccdm_b200", "csrc", cu_name)
    return _compile(host_source(cu_path), os.path.splitext(cu_name)[0])

"""The built library is Blackwell-native where DESIGN.md says so: the SASS of the default dense-layer GEMM (gemm_tma_kernel) has
tcgen05 MMAs (UTCHMMA), tensor-memory loads and stores (LDTM / STTM), TMA loads (UTMALDG) and no cp.async (LDGSTS); the two
earlier generations keep LDGSTS; the front-end kernels are CUDA-core kernels (FFMA / FADD, no tensor instructions).  CPU-only:
cuobjdump reads the in-tree .so (B200_PROFILING.md, "What proves a Blackwell-native kernel")."""
import collections
import os
import re
import shutil
import subprocess

import pytest

from conftest import ROOT

SO = os.path.join(ROOT, "streamz_b200", "lib", "libstreamz_b200.so")


@pytest.fixture(scope="module")
def sass_counts():
    if not shutil.which("cuobjdump") or not os.path.exists(SO):
        pytest.skip("needs cuobjdump and the built library")
    text = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True, check=True).stdout
    counts, cur = {}, None
    for line in text.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
        if m and cur:
            counts[cur][m.group(1)] += 1
    return counts


def _kernels(counts, needle):
    return {k: v for k, v in counts.items() if needle in k}


def test_default_gemm_uses_tcgen05_tensor_memory_and_tma(sass_counts):
    tma = _kernels(sass_counts, "gemm_tma_kernel")
    group = _kernels(sass_counts, "gemm_tma_group_kernel")
    assert len(tma) >= 12 and len(group) >= 2
    for name, c in {**tma, **group}.items():
        assert c["UTCHMMA"] >= 4, name                    # tcgen05.mma kind::tf32 (4 per k-block, x3 for 3xTF32)
        assert c["LDTM"] >= 1 and c["STTM"] >= 1, name    # accumulator read-back, A operand written to tensor memory
        assert c["UTMALDG"] == 2, name                    # one bulk tensor copy per operand and k-block
        assert c["LDGSTS"] == 0, name                     # no cp.async left in this generation
        assert c["UTCBAR"] >= 1, name                     # tcgen05.commit -> mbarrier


def test_earlier_generations_are_still_the_cp_async_kernels(sass_counts):
    for needle in ("gemm_tc_ta_kernel", "gemm_tc_async_kernel"):
        ks = _kernels(sass_counts, needle)
        assert ks
        for name, c in ks.items():
            assert c["UTCHMMA"] >= 4 and c["LDGSTS"] >= 1 and c["UTMALDG"] == 0, name


def test_front_end_kernels_are_cuda_core_kernels(sass_counts):
    for needle in ("extract_kernel", "resample_kernel"):
        ks = _kernels(sass_counts, needle)
        assert ks
        for name, c in ks.items():
            assert c["FFMA"] + c["FADD"] > 50 and c["UTCHMMA"] == 0 and c["HMMA"] == 0, name

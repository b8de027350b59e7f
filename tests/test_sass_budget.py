"""Resource budget of the built CUDA library, read from the cubins with cuobjdump (no GPU needed).

DESIGN.md argues with these numbers: the G1 accumulation kernel must fit 128 registers (4 CTAs / SM), the G2 kernel must
not spill, and the lazily reduced additions / the dedicated squaring must really remove wide multiplies from the hot loop.
A compiler or source change that silently breaks one of them shows up here instead of as a slower bench line."""
import os
import re
import shutil
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "..", "simple_zk_rollups_b200", "libzkr.so")
CUOBJDUMP = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"

pytestmark = pytest.mark.skipif(not (os.path.exists(LIB) and os.path.exists(CUOBJDUMP)),
                                reason="libzkr.so not built or cuobjdump missing")


@pytest.fixture(scope="module")
def usage():
    """mangled kernel name -> (registers, stack bytes)"""
    txt = subprocess.run([CUOBJDUMP, "--dump-resource-usage", LIB], capture_output=True, text=True, check=True).stdout
    res = {}
    for m in re.finditer(r"Function (\S+):\s*\n\s*REG:(\d+) STACK:(\d+)", txt):
        res[m.group(1)] = (int(m.group(2)), int(m.group(3)))
    assert res, "cuobjdump printed no resource usage"
    return res


def find(usage, *needles):
    hits = [k for k in usage if all(n in k for n in needles)]
    assert len(hits) == 1, (needles, hits)
    return hits[0]


def wide_multiplies(name):
    sass = subprocess.run([CUOBJDUMP, "-sass", "-fun", name, LIB], capture_output=True, text=True, check=True).stdout
    return len(re.findall(r"\bIMAD\.WIDE", sass)), len(re.findall(r"\b(?:LDL|STL)\b", sass))


def test_only_sm_100a_code_is_shipped():
    txt = subprocess.run([CUOBJDUMP, "-lelf", LIB], capture_output=True, text=True, check=True).stdout
    archs = set(re.findall(r"sm_(\d+\w?)", txt))
    assert archs == {"100a"}, archs


def test_accumulation_kernels_fit_their_occupancy(usage):
    # G1: 128 registers = 4 CTAs of 128 threads per SM (__launch_bounds__(128, 4)); a few bytes of spill are tolerated
    for variant in ("ELb1ELi0E", "ELb1ELi1E", "ELb1ELi2E"):
        regs, stack = usage[find(usage, "k_accum_affine", "FqParams", variant)]
        assert regs <= 128 and stack <= 32, (variant, regs, stack)
    # G2: two CTAs per SM, and no local memory at all in either form
    for variant in ("ELb0ELi0E", "ELb0ELi1E"):
        regs, stack = usage[find(usage, "k_accum_affine", "Fq2", variant)]
        assert regs <= 255 and stack == 0, (variant, regs, stack)


def test_lazy_forms_remove_wide_multiplies_from_the_g1_loop(usage):
    """per mixed addition: sums of products -64 wide multiplies (one reduction fewer), the two squarings -56 more
    (the kernel also holds the doubling path, so the totals are compared, not the loop alone)"""
    base, _ = wide_multiplies(find(usage, "k_accum_affine", "FqParams", "ELb1ELi0E"))
    lazy, _ = wide_multiplies(find(usage, "k_accum_affine", "FqParams", "ELb1ELi1E"))
    sqr, local = wide_multiplies(find(usage, "k_accum_affine", "FqParams", "ELb1ELi2E"))
    assert base - lazy >= 48, (base, lazy)
    assert lazy - sqr >= 48, (lazy, sqr)
    assert local <= 16, "the default G1 accumulation kernel gained local-memory traffic: %d LDL/STL" % local


def test_lazy_g2_kernel_is_smaller(usage):
    def instructions(name):
        sass = subprocess.run([CUOBJDUMP, "-sass", "-fun", name, LIB], capture_output=True, text=True, check=True).stdout
        return len(re.findall(r"^\s+/\*[0-9a-f]{4,}\*/", sass, flags=re.M))
    base = instructions(find(usage, "k_accum_affine", "Fq2", "ELb0ELi0E"))
    lazy = instructions(find(usage, "k_accum_affine", "Fq2", "ELb0ELi1E"))
    assert lazy < 0.93 * base, (base, lazy)      # measured: 11 504 -> 10 272 instructions


def test_ntt_pass_kernels_keep_two_ctas_per_sm(usage):
    names = [k for k in usage if "k_ntt_pass" in k]
    assert names
    for k in names:
        regs, stack = usage[k]
        assert regs <= 128, (k, regs)            # __launch_bounds__(256, 2)
    # the single-GPU transforms (no exchange: third template argument of NttXchg mode 0) stay nearly stack-free
    main = [k for k in names if "ILb0ELi0E" in k or "ILb1ELi0E" in k]
    assert main and all(usage[k][1] <= 32 for k in main), [(k[-60:], usage[k]) for k in main]

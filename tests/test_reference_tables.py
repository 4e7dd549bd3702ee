"""The only pieces of the reference that CAN be held against it mechanically in this container: its data tables.  Runs where
/root/reference exists (the build container) and is skipped elsewhere (the GPU box has no reference tree)."""
import re
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
REF = Path("/root/reference/src/core")
pytestmark = pytest.mark.skipif(not REF.exists(), reason="reference tree not present")


def test_sobol_blob_is_the_reference_tables(tmp_path):
    """data/sobol_tables.bin (embedded in the library, loaded by the oracle) == SOBOL_MATRICES_32 / VD_C_SOBOL_MATRICES(_INV) of
    src/core/sobolmatrices.rs, re-extracted now."""
    out = tmp_path / "sobol.bin"
    subprocess.run([sys.executable, str(ROOT / "tools" / "extract_sobol_tables.py"), str(REF / "sobolmatrices.rs"), str(out)], check=True,
                   capture_output=True)
    assert out.read_bytes() == (ROOT / "data" / "sobol_tables.bin").read_bytes()


def test_prime_tables_are_the_reference_tables(oracle):
    """PRIMES / PRIME_SUMS (src/core/lowdiscrepancy.rs:18-147) are generated, not transcribed, on our side."""
    text = (REF / "lowdiscrepancy.rs").read_text()

    def arr(name):
        i = text.index("pub const %s:" % name)
        j = text.index("[", text.index("=", i))
        k = text.index("];", j)
        return [int(t.replace("_", "")) for t in re.findall(r"[\d_]+", text[j + 1:k]) if t.strip("_")]

    primes, sums = arr("PRIMES"), arr("PRIME_SUMS")
    assert len(primes) == len(sums) == 1000
    L = oracle.load()
    assert [L.orc_prime(i) for i in range(1000)] == primes
    assert [L.orc_prime_sum(i) for i in range(1000)] == sums


def test_constants_match_the_reference_source():
    """A few literals the restatement depends on, read from the reference source."""
    pbrt = (REF / "pbrt.rs").read_text()
    assert "pub const SHADOW_EPSILON: Float = 0.0001;" in pbrt
    assert re.search(r"pub const INV_2_PI: Float = 0\.159_154_943_091_895_335_77;", pbrt)
    rng = (REF / "rng.rs").read_text()
    assert "0x853c_49e6_748f_ea9b" in rng and "0xda3e_39cb_94b9_5bdb" in rng and "0x5851_f42d_4c95_7f2d" in rng
    assert "(!b + 1) & b" in rng  # the bounded-draw threshold is restated as written
    halton = (Path("/root/reference/src/samplers") / "halton.rs").read_text()
    assert "pub const K_MAX_RESOLUTION: i32 = 128_i32;" in halton

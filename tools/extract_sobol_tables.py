#!/usr/bin/env python3
"""Extract the Sobol' generator matrices used by rs_pbrt's SobolSampler into a
compact binary blob (data/sobol_tables.bin).

These are *data*, not code: the public Joe-Kuo / Gruenschloss direction numbers
that pbrt-v3 ships as sobolmatrices.cpp and rs_pbrt transcribes in
src/core/sobolmatrices.rs (SOBOL_MATRICES_32 :7, VD_C_SOBOL_MATRICES :53463,
VD_C_SOBOL_MATRICES_INV :54155).  The reference tree only exists in the build
container, so this script is run once there and its output is committed.

Layout (little endian):
  u32 magic 'SOBL' (0x4c424f53), u32 n_dims (1024), u32 matrix_size (52),
  u32 n_vdc (25), u32 n_vdc_inv (26), u32 reserved[3]
  u32 sobol32[n_dims * matrix_size]
  u64 vdc    [n_vdc     * 52]   (row k = resolution exponent m = k+1, zero padded)
  u64 vdc_inv[n_vdc_inv * 52]
"""
import re
import struct
import sys
from pathlib import Path

SRC = Path(sys.argv[1] if len(sys.argv) > 1 else "/root/reference/src/core/sobolmatrices.rs")
DST = Path(sys.argv[2] if len(sys.argv) > 2 else Path(__file__).resolve().parent.parent / "data" / "sobol_tables.bin")

text = SRC.read_text()
num = re.compile(r"0x[0-9a-fA-F_]+")


def array_body(name):
    m = re.search(r"const %s\b[^=]*=\s*\[(.*?)\];" % re.escape(name), text, re.S)
    if m is None:
        raise SystemExit("table %s not found" % name)
    return [int(t.replace("_", "").rstrip("u"), 16) for t in num.findall(m.group(1))]


def clean(vals):
    return vals


sobol32 = array_body("SOBOL_MATRICES_32")
assert len(sobol32) == 1024 * 52, len(sobol32)
vdc = [array_body("M%d" % k) for k in range(1, 26)]
vdc_inv = [array_body("MI%d" % k) for k in range(1, 27)]
for k, a in enumerate(vdc, 1):
    assert len(a) == 52 - 2 * k, (k, len(a))
for k, a in enumerate(vdc_inv, 1):
    assert len(a) == 2 * k, (k, len(a))

out = bytearray()
out += struct.pack("<8I", 0x4C424F53, 1024, 52, 25, 26, 0, 0, 0)
out += struct.pack("<%dI" % len(sobol32), *sobol32)
for a in vdc:
    out += struct.pack("<52Q", *(a + [0] * (52 - len(a))))
for a in vdc_inv:
    out += struct.pack("<52Q", *(a + [0] * (52 - len(a))))
DST.parent.mkdir(parents=True, exist_ok=True)
DST.write_bytes(bytes(out))
print("wrote", DST, len(out), "bytes")

"""GPU parity of the data-parallel LZ stage (NAFGPU_LZ=shared: k_zlc_find / k_zlc_define / k_zlc_finish, zstd_enc.cu).  The switch
is read once per process, so the checks run in a process of their own (tools/zlc_gpu_check.py): frames of single streams equal the
CPU emulation of the same HD bodies byte for byte (tests/emu/emu_zlzc.cpp, itself pinned to libzstd / the oracle / the serial
restatement by tests/test_emu_zenc.py), and whole files decode back to the text on the oracle and on the device."""
import json
import os
import subprocess
import sys

import pytest

import helpers

pytestmark = pytest.mark.gpu


def test_shared_table_lz_frames_equal_cpu_emulation_and_decode(gpu, tmp_path):
    env = dict(os.environ, TMPDIR=str(tmp_path))
    p = subprocess.run([sys.executable, os.path.join(helpers.ROOT, "tools", "zlc_gpu_check.py"), "0"], capture_output=True, text=True, env=env, timeout=600)
    lines = [json.loads(x) for x in p.stdout.splitlines() if x.startswith("{")]
    bad = [x for x in lines if x.get("step") in ("A", "B") and not all(v for k, v in x.items() if k in ("equals_emulation", "device_decodes", "oracle_decodes"))]
    assert p.returncode == 0 and not bad, (bad, p.stderr[-2000:])
    verdict = [x for x in lines if x.get("step") == "verdict"]
    assert verdict and verdict[0]["all_ok"]
    assert sum(x.get("step") == "A" for x in lines) >= 15 and sum(x.get("step") == "B" for x in lines) == 2
    fq = [x for x in lines if x.get("step") == "B" and x["case"] == "fastq"][0]
    assert fq["ratio"] < 0.365                                  # level 1 (entropy only) gives 0.381 on this shape, `ennaf -1` 0.358

"""GPU: the drop-in command-line tools (bin/ennaf, bin/unnaf over libnafgpu.so) run the reference's own
perl-suite cases (tests/{alphabet,charcount,small,large}/*.test): `ennaf ARGS in.fa | unnaf ARGS` must print the
pinned stdout, and ennaf's stderr must equal the pinned unexpected-character report."""
import os
import subprocess

import pytest

import helpers

pytestmark = pytest.mark.gpu
BIN = os.path.join(helpers.ROOT, "bin")


def run(tool, args, stdin=b""):
    p = subprocess.run([os.path.join(BIN, tool), *args], input=stdin, capture_output=True)
    return p.returncode, p.stdout, p.stderr


@pytest.fixture(scope="module", autouse=True)
def built():
    if not all(os.access(os.path.join(BIN, t), os.X_OK) for t in ("ennaf", "unnaf")):
        import __graft_entry__
        __graft_entry__.build()


def test_reference_suite_through_cli():
    # every process pays ~1.5 s of CUDA context creation: run every 4th case here, all 60 run through the
    # library in test_gpu_encode.py::test_encode_reference_suite
    for case in helpers.manifest("ref_suite")[::4]:
        src = os.path.join(helpers.GOLDEN, "ref_suite", case["input"])
        rc, naf, eerr = run("ennaf", [*case["ennaf_args"], src, "-c"])
        assert rc == 0, (case["name"], eerr)
        assert eerr == helpers.golden("ref_suite", case["set"], case["name"] + ".e.err"), case["name"]
        rc, out, uerr = run("unnaf", case["unnaf_args"], naf)
        assert rc == 0, (case["name"], uerr)
        assert out == helpers.golden("ref_suite", case["set"], case["name"] + ".out"), case["name"]
        assert uerr == helpers.golden("ref_suite", case["set"], case["name"] + ".u.err"), case["name"]


def test_cli_views_match_reference():
    """a CUDA context per process makes each invocation ~1 s: a representative subset of cases x views"""
    picked = {"fq_small": ["fastq", "fasta", "ids", "lengths", "charcount --no-mask", "format", "sizes"],
              "mask_runs": ["fasta", "fasta --line-length 33", "mask", "total-mask-length", "seq --no-mask", "sequences", "4bit"],
              "protein": ["fasta --no-mask", "names", "number", "part-list", "4bit"],
              "title_linelen": ["title", "fasta", "total-length"]}
    for case in helpers.manifest("cases"):
        if case["name"] not in picked:
            continue
        naf = helpers.golden("cases", case["name"] + ".naf")
        for key in picked[case["name"]]:
            exp = case["views"].get(key)
            if exp is None:                       # "sizes" is not in the manifest (compressed sizes differ by design)
                rc, out, uerr = run("unnaf", ["--" + key], naf)
                assert rc == 0 and out.startswith(b"IDs: "), (case["name"], key, uerr)
                continue
            parts = key.split()
            rc, out, uerr = run("unnaf", ["--" + parts[0], *parts[1:]], naf)
            assert (rc != 0) == (exp["rc"] != 0), (case["name"], key, uerr)
            if rc == 0:
                assert (len(out), helpers.sha(out)) == (exp["size"], exp["sha256"]), (case["name"], key)
            else:
                assert uerr.decode("latin-1") == exp["stderr"], (case["name"], key)


def test_cli_files_and_errors(tmp_path):
    fa = tmp_path / "x.fa"
    fa.write_bytes(b">a b\nACGTNacgt\n>c\nGG\n")
    naf = tmp_path / "x.naf"
    rc, _, e = run("ennaf", [str(fa), "-o", str(naf)])
    assert rc == 0 and naf.exists(), e
    out = tmp_path / "y.fa"
    rc, _, e = run("unnaf", [str(naf), "-o", str(out)])
    assert rc == 0 and out.read_bytes() == fa.read_bytes(), e
    rc, o, e = run("unnaf", ["--fastq", str(naf)])
    assert rc == 1 and e == b"unnaf error: FASTQ output requested, but input has no qualities\n"
    rc, o, e = run("ennaf", ["-c"], b"ACGT\n")
    assert rc == 1 and e == b"ennaf error: input data is in unknown format - first non-space character is neither '>' nor '@'\n"
    rc, o, e = run("unnaf", [], b"not a naf file")
    assert rc == 1 and e == b"unnaf error: not a NAF format\n"
    rc, o, e = run("ennaf", ["--bogus"])
    assert rc == 1 and e == b'ennaf error: unknown or incomplete argument "--bogus"\n'

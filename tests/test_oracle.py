"""not-gpu: pin the CPU oracle against the reference's own golden vectors and reference-made files."""
import pytest

import helpers


def test_vle_known_answers(oracle):
    """NAFv2.pdf section 10 examples"""
    import ctypes as C
    for v, enc in [(0, b"\x00"), (128, b"\x81\x00"), (34359738368, b"\x81\x80\x80\x80\x80\x00"), (127, b"\x7f"), (150, b"\x81\x16")]:
        ob = helpers.OBuf()
        oracle.lib.onaf_put_vle(C.byref(ob), v)
        assert C.string_at(ob.data, ob.size) == enc
        oracle.lib.obuf_free(C.byref(ob))
        pos, val, err = C.c_size_t(0), C.c_uint64(), C.create_string_buffer(256)
        assert oracle.lib.onaf_get_vle(enc, len(enc), C.byref(pos), C.byref(val), err) == 0
        assert val.value == v and pos.value == len(enc)
        from naf_b200 import container
        assert container.put_vle(v) == enc and container.get_vle(enc, 0) == (v, len(enc))


def test_oracle_zstd_vs_golden_frames(oracle):
    for e in helpers.manifest("zstd"):
        out = oracle.zstd_decompress(helpers.golden("zstd", e["frame"]))
        assert helpers.sha(out) == e["sha256"], e["frame"]


def test_oracle_reference_suite(oracle):
    """the 60 cases of the reference's perl suite: oracle ennaf -> oracle unnaf reproduces the pinned stdout and
    the pinned ennaf stderr (unexpected-character report); oracle unnaf on the reference-made .naf too."""
    for case in helpers.manifest("ref_suite"):
        text = helpers.golden("ref_suite", case["input"])
        ekw, ukw = helpers.parse_ennaf_args(case["ennaf_args"]), helpers.parse_unnaf_args(case["unnaf_args"])
        expect = helpers.golden("ref_suite", case["set"], case["name"] + ".out")
        naf, report = oracle.encode(text, **ekw)
        assert report == helpers.golden("ref_suite", case["set"], case["name"] + ".e.err"), case["name"]
        assert oracle.decode(naf, **ukw) == expect, case["name"]
        assert oracle.decode(helpers.golden("ref_suite", case["naf"]), **ukw) == expect, case["name"]


def test_oracle_cases_all_views(oracle):
    for case in helpers.manifest("cases"):
        naf = helpers.golden("cases", case["name"] + ".naf")
        text = helpers.golden("cases", case["name"] + ".txt.gz")
        assert helpers.sha(text) == case["text_sha256"]
        for key, exp in case["views"].items():
            parts = key.split()
            kw = helpers.parse_unnaf_args(["--" + parts[0]] + parts[1:])
            if exp["rc"] != 0:
                with pytest.raises(ValueError):
                    oracle.decode(naf, **kw)
                continue
            got = oracle.decode(naf, **kw)
            assert (len(got), helpers.sha(got)) == (exp["size"], exp["sha256"]), (case["name"], key)
        # encode side: the oracle's streams, stored raw, decode to the same text through the oracle
        ekw = helpers.parse_ennaf_args(case["ennaf_args"])
        mine, report = oracle.encode(text, **ekw)
        assert report.decode("latin-1") == case["ennaf_stderr"]
        for key in ("fasta", "fastq", "seq", "4bit", "ids", "names", "lengths", "mask"):
            if case["views"][key]["rc"] == 0:
                assert helpers.sha(oracle.decode(mine, key)) == case["views"][key]["sha256"], (case["name"], key)


@pytest.mark.skipif(not helpers.have_ref(), reason="oracle/_ref binaries not built")
def test_oracle_encode_is_decodable_by_reference(oracle, tmp_path):
    from naf_b200 import synth
    text = synth.fastq(2000, 150, seed=5)
    naf, _ = oracle.encode(text)
    rc, out, err = helpers.ref_run("unnaf", [], naf)
    assert rc == 0 and out == text
    rc, refnaf, err = helpers.ref_run("ennaf", ["-c"], text, tmp=str(tmp_path))
    assert oracle.decode(refnaf) == text


def test_config1_roundtrip_oracle(oracle):
    """BASELINE config 1: 1k x 150 bp ACGT-only FASTA, bit-exact round trip on the CPU"""
    from naf_b200 import synth
    text = synth.fasta_reads(1000, 150, seed=42)
    naf, report = oracle.encode(text)
    assert report == b"" and oracle.decode(naf) == text

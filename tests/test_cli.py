"""`bigseqkit <cmd>` CLI surface (bigseqkit_b200/cli.py): flags -> the reference option JSON -> libbsk.so -> files."""
import os

import pytest

import oracle
from bigseqkit_b200 import api, cli
from cases import FA_SIMPLE, FQ_SIMPLE


def test_flag_surface_matches_reference_names():
    p = cli.build_parser()
    a = p.parse_args(["seq", "-r", "-p", "-w", "0", "-m", "10", "x.fq"])
    name, o = cli.options(a)
    assert name == "SeqTransform" and o["Reverse"] and o["Complement"] and o["MinLen"] == 10 and o["Config"]["LineWidth"] == 0
    assert o["GapLetters"] == "- \t." and o["MaxLen"] == -1 and o["QualAsciiBase"] == 33  # bigseqkit/seq.go:32-55
    a = p.parse_args(["translate", "-f", "6", "-T", "11", "--trim", "x.fa"])
    assert cli.options(a)[1]["Frame"] == ["6"] and cli.options(a)[1]["TranslTable"] == 11
    a = p.parse_args(["locate", "-p", "ACG,TTT", "-p", "GGA", "-P", "--bed", "x.fa"])
    assert cli.options(a)[1]["Pattern"] == ["ACG", "TTT", "GGA"] and cli.options(a)[1]["Bed"]
    a = p.parse_args(["stats", "-a", "-T", "x.fa"])
    assert cli.options(a)[1] == {"Config": cli._config(a), "Tabular": True, "GapLetters": "- .", "All": True, "FqEncoding": "sanger"}
    a = p.parse_args(["rmdup", "-s", "-i", "x.fq"])
    assert cli.options(a)[0] == "RmDup" and cli.options(a)[1]["BySeq"] and cli.options(a)[1]["IgnoreCase"]
    a = p.parse_args(["grep", "-s", "-p", "GTA", "-v", "x.fa"])
    assert cli.options(a)[1]["BySeq"] and cli.options(a)[1]["InvertMatch"]
    a = p.parse_args(["subseq", "-r", "2:-2", "x.fa"])
    assert cli.options(a)[1]["Region"] == "2:-2"


def _run_cli(lib, monkeypatch, argv):
    monkeypatch.setattr(api, "_default", lib)  # the CLI takes the default library; tests point it at the emulator / GPU
    return cli.main(argv)


def test_cli_end_to_end(lib, monkeypatch, tmp_path, capsys):
    fq = tmp_path / "r.fq"
    fq.write_bytes(FQ_SIMPLE)
    fa = tmp_path / "s.fa"
    fa.write_bytes(FA_SIMPLE)
    assert _run_cli(lib, monkeypatch, ["seq", "-r", "-p", str(fq), "-o", str(tmp_path / "o1")]) == 0
    assert (tmp_path / "o1").read_bytes() == oracle.seq(FQ_SIMPLE, {"Reverse": True, "Complement": True})[0]
    assert _run_cli(lib, monkeypatch, ["seq", "-w", "4", str(fa)]) == 0  # default out name: <input>-out
    assert (tmp_path / "s.fa-out").read_bytes() == oracle.seq(FA_SIMPLE, {"Config": {"LineWidth": 4}})[0]
    assert _run_cli(lib, monkeypatch, ["translate", "-f", "6", "-x", str(fa), "-o", str(tmp_path / "o2")]) == 0
    assert (tmp_path / "o2").read_bytes() == oracle.translate(FA_SIMPLE, {"Frame": ["6"], "AllowUnknownCodon": True})[0]
    assert _run_cli(lib, monkeypatch, ["rmdup", "-s", str(fq), str(fq), "-o", str(tmp_path / "o3")]) == 0
    assert (tmp_path / "o3").read_bytes().count(b"@") >= 2
    # rmdup -d / -D: removed records and "count\tids" rows land in the named files (flag meaning)
    twice = tmp_path / "twice.fq"
    twice.write_bytes(FQ_SIMPLE * 2)
    assert _run_cli(lib, monkeypatch, ["rmdup", "-s", "-d", str(tmp_path / "d.fq"), "-D", str(tmp_path / "d.txt"), str(twice),
                                       "-o", str(tmp_path / "o5")]) == 0
    assert (tmp_path / "o5").read_bytes() == oracle.rmdup(FQ_SIMPLE * 2, {"BySeq": True})[0]
    assert ((tmp_path / "d.fq").read_bytes(), (tmp_path / "d.txt").read_bytes()) == oracle.rmdup_dups(FQ_SIMPLE * 2, {"BySeq": True})
    assert (tmp_path / "d.txt").read_bytes().startswith(b"2\t")
    assert _run_cli(lib, monkeypatch, ["fq2fa", str(fq), "-o", str(tmp_path / "o6")]) == 0
    assert (tmp_path / "o6").read_bytes() == oracle.fq2fa(FQ_SIMPLE, {})[0]
    capsys.readouterr()
    assert _run_cli(lib, monkeypatch, ["stats", "-T", str(fa), str(fq)]) == 0
    out = capsys.readouterr().out.split("\n")
    assert out[0].startswith("file\tformat\ttype") and out[1].startswith("input0\tN/A\t") and out[2].startswith("input1\tN/A\t")
    assert out[1] == oracle.stats(FA_SIMPLE, {"Tabular": True})[1].split("\n")[1]
    # a flag error surfaces with the reference's text and a non-zero status
    assert _run_cli(lib, monkeypatch, ["seq", "-l", "-u", str(fa), "-o", str(tmp_path / "o4")]) == 1
    assert "could not give both flags -l (--lower-case) and -u (--upper-case)" in capsys.readouterr().err


def test_cli_partitions_union_and_new_commands(lib, monkeypatch, tmp_path, capsys):
    from bigseqkit_b200 import synth
    fq = synth.fastq_reads(60 << 10, seed=41, dup_frac=0.3).tobytes()
    a, b = tmp_path / "a.fq", tmp_path / "b.fq"
    a.write_bytes(fq)
    b.write_bytes(fq[: len(fq) // 2 and oracle.frame(fq)[len(oracle.frame(fq)) // 2]])
    both = a.read_bytes() + b.read_bytes()
    # rmdup over several inputs dedups the Union of them (bigseqkit-cli/helper.go:131-138, bigseqkit/rmdup.go:97)
    assert _run_cli(lib, monkeypatch, ["rmdup", "-s", str(a), str(b), "-o", str(tmp_path / "u.fq")]) == 0
    assert (tmp_path / "u.fq").read_bytes() == oracle.rmdup(both, {"BySeq": True})[0]
    # ... and over the partitions of one input
    assert _run_cli(lib, monkeypatch, ["rmdup", "-s", "--partitions", "3", str(a), "-o", str(tmp_path / "u3.fq")]) == 0
    assert (tmp_path / "u3.fq").read_bytes() == oracle.rmdup(fq, {"BySeq": True})[0]
    # partitions + --merge: one file in partition order; without --merge: a directory of parts (StoreFASTXN)
    exp = oracle.seq(fq, {"Reverse": True, "Complement": True})[0]
    assert _run_cli(lib, monkeypatch, ["seq", "-r", "-p", "--partitions", "4", "--merge", str(a), "-o", str(tmp_path / "m.fq")]) == 0
    assert (tmp_path / "m.fq").read_bytes() == exp
    assert _run_cli(lib, monkeypatch, ["seq", "-r", "-p", "--partitions", "4", str(a), "-o", str(tmp_path / "parts")]) == 0
    names = sorted(os.listdir(tmp_path / "parts"))
    assert names == ["part%05d" % k for k in range(4)]
    assert b"".join((tmp_path / "parts" / n).read_bytes() for n in names) == exp
    # stats over partitions: histograms merged with sum semantics
    capsys.readouterr()
    assert _run_cli(lib, monkeypatch, ["stats", "-T", "-a", "--partitions", "3", str(a)]) == 0
    assert capsys.readouterr().out == oracle.stats(fq, {"Tabular": True, "All": True})[1]
    # duplicate / head / range
    assert _run_cli(lib, monkeypatch, ["duplicate", "-n", "3", str(a), "-o", str(tmp_path / "d3")]) == 0
    assert (tmp_path / "d3").read_bytes() == oracle.duplicate(fq, 3)[0]
    assert _run_cli(lib, monkeypatch, ["head", "-n", "7", "--partitions", "2", str(a), str(b), "-o", str(tmp_path / "h7")]) == 0
    assert (tmp_path / "h7").read_bytes() == oracle.head(both, 7)[0]
    assert _run_cli(lib, monkeypatch, ["range", "-r", "5:9", str(a), "-o", str(tmp_path / "r59")]) == 0
    assert (tmp_path / "r59").read_bytes() == oracle.range_(fq, 4, 9)[0]

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

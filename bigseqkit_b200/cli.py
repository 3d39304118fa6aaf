"""`bigseqkit <cmd> [flags] files...` -- the reference CLI surface (bigseqkit-cli/*.go, cobra) for the accelerated
commands, on top of libbsk.so.  Flag names, shorthands and defaults follow bigseqkit-cli/seq.go:54-73,
stats.go:61-65, rmdup.go:45-51, locate.go:61-75, grep.go:81-96, subseq.go:56-67, translate.go:86-94, duplicate.go,
range.go, head.go and the persistent flags of bigseqkit-cli/helper.go:161-173.  Input type is sniffed like
helper.go:47-85; `stats` prints its table to stdout (bigseqkit-cli/stats.go:10-28).

Partitions and devices.  Every input file is cut into `--partitions` record-aligned ranges (bsk_shard_bounds; default
one per file) and the ranges are handed round-robin to the CUDA devices of `--devices` (one ctx and one host thread
each); every range streams file -> HBM -> file through bsk_run_file.  Outputs go to `-o` (default `<first input>-out`,
helper.go:108-121): one file in partition order with --merge or when there is a single partition (StoreFASTX), else a
directory of `part<k>` files (StoreFASTXN).  `rmdup` and `range` / `head` see all partitions as ONE dataframe, in input
order (the reference Unions the inputs, helper.go:131-138): they run on the first device with bsk_set_union.

    python -m bigseqkit_b200.cli seq -r -p reads.fq -o out.fq
    python -m bigseqkit_b200.cli translate -f 6 --partitions 8 --devices 0,1,2,3 --merge cds.fa -o prot.fa
"""
import argparse
import os
import shutil
import sys
import threading

from .api import BskError, Operator, shard_bounds


def _persistent(p):
    p.add_argument("-t", "--seq-type", default="auto")
    p.add_argument("-w", "--line-width", type=int, default=60)
    p.add_argument("--id-regexp", default=r"^(\S+)\s?")
    p.add_argument("--id-ncbi", action="store_true")
    p.add_argument("-o", "--out-file", default="")
    p.add_argument("--quiet", action="store_true")
    p.add_argument("--alphabet-guess-seq-length", type=int, default=10000)
    p.add_argument("--infile-list", default="")
    p.add_argument("--merge", action="store_true")
    p.add_argument("--partitions", type=int, default=0)
    p.add_argument("--order", action="store_true")
    p.add_argument("--device", type=int, default=0, help="CUDA device (not in the reference: executors pick one each)")
    p.add_argument("--devices", default="", help="comma list of CUDA devices the partitions are spread over (overrides --device)")
    p.add_argument("files", nargs="*")


def build_parser():
    ap = argparse.ArgumentParser(prog="bigseqkit", description="BigSeqKit per-record commands on a B200 (libbsk.so)")
    sub = ap.add_subparsers(dest="cmd", required=True)

    p = sub.add_parser("seq", help="transform sequences (extract ID, filter by length, remove gaps, reverse complement...)")
    for s, l in (("-r", "--reverse"), ("-p", "--complement"), ("-n", "--name"), ("-s", "--seq"), ("-q", "--qual"),
                 ("-i", "--only-id"), ("-g", "--remove-gaps"), ("-l", "--lower-case"), ("-u", "--upper-case"),
                 ("-k", "--color"), ("-v", "--validate-seq")):
        p.add_argument(s, l, action="store_true")
    p.add_argument("--dna2rna", action="store_true")
    p.add_argument("--rna2dna", action="store_true")
    p.add_argument("-G", "--gap-letters", default="- \t.")
    p.add_argument("-V", "--validate-seq-length", type=int, default=10000)
    p.add_argument("-m", "--min-len", type=int, default=-1)
    p.add_argument("-M", "--max-len", type=int, default=-1)
    p.add_argument("-b", "--qual-ascii-base", type=int, default=33)
    p.add_argument("-Q", "--min-qual", type=float, default=-1)
    p.add_argument("-R", "--max-qual", type=float, default=-1)
    _persistent(p)

    p = sub.add_parser("stats", help="simple statistics of FASTA/Q files")
    p.add_argument("-T", "--tabular", action="store_true")
    p.add_argument("-G", "--gap-letters", default="- .")
    p.add_argument("-a", "--all", action="store_true")
    p.add_argument("-e", "--skip-err", action="store_true")
    p.add_argument("-E", "--fq-encoding", default="sanger")
    p.add_argument("-b", "--basename", action="store_true")
    p.add_argument("-i", "--stdin-label", default="-")
    _persistent(p)

    p = sub.add_parser("rmdup", help="remove duplicated sequences by id/name/sequence")
    p.add_argument("-n", "--by-name", action="store_true")
    p.add_argument("-s", "--by-seq", action="store_true")
    p.add_argument("-i", "--ignore-case", action="store_true")
    p.add_argument("-d", "--dup-seqs-file", default="")
    p.add_argument("-D", "--dup-num-file", default="")
    p.add_argument("-r", "--consider-revcom", action="store_true")
    p.add_argument("-P", "--only-positive-strand", action="store_true")
    _persistent(p)

    p = sub.add_parser("locate", help="locate subsequences/motifs")
    p.add_argument("-p", "--pattern", action="append", default=[])
    p.add_argument("-f", "--pattern-file", default="")
    for s, l in (("-d", "--degenerate"), ("-r", "--use-regexp"), ("-F", "--use-fmi"), ("-i", "--ignore-case"),
                 ("-P", "--only-positive-strand"), ("-G", "--non-greedy"), ("-M", "--hide-matched"), ("-c", "--circular"),
                 ("-I", "--immediate-output")):
        p.add_argument(s, l, action="store_true")
    p.add_argument("--gtf", action="store_true")
    p.add_argument("--bed", action="store_true")
    p.add_argument("-V", "--validate-seq-length", type=int, default=10000)
    p.add_argument("-m", "--max-mismatch", type=int, default=0)
    _persistent(p)

    p = sub.add_parser("grep", help="search sequences by ID/name/sequence/sequence motifs")
    p.add_argument("-p", "--pattern", action="append", default=[])
    p.add_argument("-f", "--pattern-file", default="")
    for s, l in (("-r", "--use-regexp"), ("-v", "--invert-match"), ("-n", "--by-name"), ("-s", "--by-seq"),
                 ("-P", "--only-positive-strand"), ("-i", "--ignore-case"), ("-d", "--degenerate"), ("-c", "--circular"),
                 ("-I", "--immediate-output"), ("-C", "--count")):
        p.add_argument(s, l, action="store_true")
    p.add_argument("--delete-matched", action="store_true")
    p.add_argument("-m", "--max-mismatch", type=int, default=0)
    p.add_argument("-R", "--region", default="")
    _persistent(p)

    p = sub.add_parser("subseq", help="get subsequences by region")
    p.add_argument("--chr", action="append", default=[])
    p.add_argument("-r", "--region", default="")
    p.add_argument("--gtf", default="")
    p.add_argument("--feature", action="append", default=[])
    p.add_argument("-u", "--up-stream", type=int, default=0)
    p.add_argument("-d", "--down-stream", type=int, default=0)
    p.add_argument("-f", "--only-flank", action="store_true")
    p.add_argument("--bed", default="")
    p.add_argument("--gtf-tag", default="gene_id")
    _persistent(p)

    p = sub.add_parser("fq2fa", help="convert FASTQ to FASTA")  # bigseqkit-cli/fq2fa.go:27-28
    _persistent(p)

    p = sub.add_parser("duplicate", help="duplicate sequences N times")  # bigseqkit-cli/duplicate.go
    p.add_argument("-n", "--times", type=int, default=1)
    _persistent(p)

    p = sub.add_parser("range", help="print FASTA/Q records in a range (start:end)")  # bigseqkit-cli/range.go
    p.add_argument("-r", "--range", default="")
    _persistent(p)

    p = sub.add_parser("head", help="print first N FASTA/Q records")  # bigseqkit-cli/head.go
    p.add_argument("-n", "--number", type=int, default=10)
    _persistent(p)

    p = sub.add_parser("translate", help="translate DNA/RNA to protein sequence (supporting ambiguous bases)")
    p.add_argument("-T", "--transl-table", type=int, default=1)
    p.add_argument("-f", "--frame", action="append", default=[])
    p.add_argument("--trim", action="store_true")
    p.add_argument("--clean", action="store_true")
    p.add_argument("-x", "--allow-unknown-codon", action="store_true")
    p.add_argument("-M", "--init-codon-as-M", action="store_true")
    p.add_argument("-l", "--list-transl-table", type=int, default=-1)
    p.add_argument("-L", "--list-transl-table-with-amb-codons", type=int, default=-1)
    p.add_argument("-F", "--append-frame", action="store_true")
    _persistent(p)
    return ap


def _split_csv(values):
    """cobra StringSlice: repeated flags and comma-separated values"""
    out = []
    for v in values:
        out.extend(x for x in v.split(",") if x != "")
    return out


def _config(a):
    return {"SeqType": a.seq_type, "LineWidth": a.line_width, "IDRegexp": a.id_regexp, "IDNCBI": a.id_ncbi, "Quiet": a.quiet,
            "AlphabetGuessSeqLength": a.alphabet_guess_seq_length}


def options(a):
    """flags -> (operator name, reference option JSON) as the SeqKit<X>Options builders of bigseqkit/*.go do"""
    cfg = _config(a)
    c = a.cmd
    if c == "seq":
        return "SeqTransform", {
            "Config": cfg, "Reverse": a.reverse, "Complement": a.complement, "Name": a.name, "Seq": a.seq, "Qual": a.qual,
            "OnlyId": a.only_id, "RemoveGaps": a.remove_gaps, "GapLetters": a.gap_letters, "LowerCase": a.lower_case,
            "UpperCase": a.upper_case, "Dna2rna": a.dna2rna, "Rna2dna": a.rna2dna, "ValidateSeq": a.validate_seq,
            "ValidateSeqLength": a.validate_seq_length, "MinLen": a.min_len, "MaxLen": a.max_len,
            "QualAsciiBase": a.qual_ascii_base, "MinQual": a.min_qual, "MaxQual": a.max_qual}
    if c == "stats":  # the reference CLI forgets to forward -G / -E; they are forwarded here
        return "Stats", {"Config": cfg, "Tabular": a.tabular, "GapLetters": a.gap_letters, "All": a.all,
                         "FqEncoding": a.fq_encoding}
    if c == "rmdup":
        return "RmDup", {"Config": cfg, "ByName": a.by_name, "BySeq": a.by_seq, "IgnoreCase": a.ignore_case,
                         "DupSeqsFile": a.dup_seqs_file, "DupNumFile": a.dup_num_file,
                         "OnlyPositiveStrand": a.only_positive_strand}
    if c == "locate":
        o = {"Config": cfg, "Pattern": _split_csv(a.pattern), "PatternFile": a.pattern_file, "Degenerate": a.degenerate,
             "UseRegexp": a.use_regexp, "UseFmi": a.use_fmi, "IgnoreCase": a.ignore_case,
             "OnlyPositiveStrand": a.only_positive_strand, "ValidateSeqLength": a.validate_seq_length,
             "NonGreedy": a.non_greedy, "Gtf": a.gtf, "Bed": a.bed, "MaxMismatch": a.max_mismatch,
             "HideMatched": a.hide_matched, "Circular": a.circular}
        return "Locate", o
    if c == "grep":
        pats = _split_csv(a.pattern)
        if a.pattern_file:  # one pattern per line, empty lines skipped (bigseqkit-lib/grep.go:124,199)
            pats = [ln for ln in open(a.pattern_file).read().split("\n") if ln != ""]
        return "Grep", {"Config": cfg, "Pattern": pats, "UseRegexp": a.use_regexp, "DeleteMatched": a.delete_matched,
                        "InvertMatch": a.invert_match, "ByName": a.by_name, "BySeq": a.by_seq,
                        "OnlyPositiveStrand": a.only_positive_strand, "MaxMismatch": a.max_mismatch,
                        "IgnoreCase": a.ignore_case, "Degenerate": a.degenerate, "Region": a.region,
                        "Circular": a.circular, "Count": a.count}
    if c == "subseq":
        return "SubseqTransform", {"Config": cfg, "Chr": a.chr, "Region": a.region, "Gtf": a.gtf, "Feature": a.feature,
                                   "UpStream": a.up_stream, "DownStream": a.down_stream, "OnlyFlank": a.only_flank,
                                   "Bed": a.bed, "GtfTag": a.gtf_tag}
    if c == "fq2fa":
        return "Fq2Fa", {"Config": cfg}
    if c == "duplicate":
        return "Duplicate", {"Config": cfg, "Times": a.times}
    if c == "head":  # bigseqkit/head.go:33-44: Range "1:N"
        return "Range", {"Config": cfg, "Start": 0, "End": a.number}
    if c == "range":
        return "Range", dict(_parse_range(a.range), Config=cfg)
    if c == "translate":
        return "Translate", {"Config": cfg, "TranslTable": a.transl_table, "Frame": _split_csv(a.frame) or ["1"],
                             "Trim": a.trim, "Clean": a.clean, "AllowUnknownCodon": a.allow_unknown_codon,
                             "InitCodonAsM": a.init_codon_as_M, "ListTranslTable": a.list_transl_table,
                             "ListTranslTableWithAmbCodons": a.list_transl_table_with_amb_codons,
                             "AppendFrame": a.append_frame}
    raise SystemExit("unknown command " + c)


def _parse_range(r):
    """-r start:end, 1-based and inclusive like seqkit range (bigseqkit/range.go:46-76).  The snapshot's driver then
    rejects every start <= end (range.go:85), so the command cannot run there; the intended reading is used here:
    records start .. end.  Negative bounds (counted from the end) need the record count and are not supported."""
    if r == "":
        raise SystemExit("flag -r (--range) needed")
    parts = r.split(":")
    start = int(parts[0])
    end = int(parts[1]) if len(parts) > 1 else -1
    if start == 0 or end == 0:
        raise SystemExit("either start and end should not be 0")
    if start < 0 or end < -1:
        raise SystemExit("negative range bounds are not supported")
    return {"Start": start - 1, "End": (1 << 62) if end == -1 else end}


def input_files(a):
    files = list(a.files)
    if a.infile_list:
        files += [ln.strip() for ln in open(a.infile_list) if ln.strip()]
    if not files:
        raise SystemExit("bigseqkit: no input files")
    return files


def _devices(a):
    if a.devices:
        return [int(x) for x in a.devices.split(",") if x != ""]
    return [a.device]


def _plan(files, partitions):
    """[(file, offset, length)] in input order: every file cut into `partitions` record-aligned ranges"""
    plan = []
    for f in files:
        if partitions and partitions > 1:
            b = shard_bounds(f, partitions)
            plan += [(f, b[i], b[i + 1] - b[i]) for i in range(partitions) if b[i + 1] > b[i] or i == 0]
        else:
            plan.append((f, 0, 0))
    return plan


def main(argv=None):
    a = build_parser().parse_args(argv)
    op_name, opts = options(a)
    files = input_files(a)
    devices = _devices(a)
    try:
        if a.cmd == "stats":  # one header + one row per input, file label input<i>, format label N/A (cli/stats.go:10-25)
            rows = []
            for i, f in enumerate(files):
                parts = _plan([f], a.partitions)
                ops = [Operator("Stats", opts, device=d) for d in devices[:len(parts)]]
                try:
                    def work(k, op):
                        for j in range(k, len(parts), len(ops)):
                            op.call_file(parts[j][0], parts[j][1], parts[j][2], partition_id=j)
                    _run_threads(work, ops)
                    for other in ops[1:]:  # StatsReduce: sum semantics (bigseqkit-lib/stats.go:119-137)
                        ops[0].stats_merge(other)
                    rows.append(ops[0].stats_render("input%d" % i, "N/A"))
                finally:
                    for op in ops:
                        op.close()
            out = rows[0] + "".join(r.split("\n", 1)[1] for r in rows[1:]) if a.tabular else "".join(rows)
            sys.stdout.write(out)
            return 0
        out_path = a.out_file or (files[0] + "-out" if len(files) == 1 else None)
        if out_path is None:
            raise SystemExit("out file -o required")
        plan = _plan(files, a.partitions)
        one_dataframe = a.cmd in ("rmdup", "range", "head")
        single = len(plan) == 1 or one_dataframe
        if single:  # one ctx, partitions in order, output appended as it is produced
            open(out_path, "wb").close()
            off = 0
            with Operator(op_name, opts, device=devices[0]) as op:
                op.set_elem_offsets(False)  # the merged file needs no element table
                op.set_union(one_dataframe)
                for i, (f, o, n) in enumerate(plan):
                    off += op.call_file(f, o, n, out_path, off, partition_id=i)[0]
                if a.cmd == "rmdup":  # bigseqkit-lib/rmdup.go:245-275
                    if a.dup_seqs_file:
                        open(a.dup_seqs_file, "wb").write(op.rmdup_dup_seqs())
                    if a.dup_num_file:
                        open(a.dup_num_file, "wb").write(op.rmdup_dup_num())
            return 0
        # several independent partitions: one part file each, spread over the devices
        part_dir = out_path + ".parts" if a.merge else out_path
        if os.path.isdir(part_dir):
            shutil.rmtree(part_dir)
        elif os.path.exists(part_dir):
            os.remove(part_dir)
        os.makedirs(part_dir)
        names = [os.path.join(part_dir, "part%05d" % k) for k in range(len(plan))]
        ops = [Operator(op_name, opts, device=d) for d in devices[:len(plan)]]
        try:
            def work(k, op):
                op.set_elem_offsets(False)
                for j in range(k, len(plan), len(ops)):
                    open(names[j], "wb").close()
                    op.call_file(plan[j][0], plan[j][1], plan[j][2], names[j], 0, partition_id=j)
            _run_threads(work, ops)
        finally:
            for op in ops:
                op.close()
        if a.merge:
            with open(out_path, "wb") as dst:
                for nm in names:
                    with open(nm, "rb") as src:
                        shutil.copyfileobj(src, dst, 16 << 20)
            shutil.rmtree(part_dir)
    except BskError as e:
        sys.stderr.write("bigseqkit: %s\n" % e)
        return 1
    return 0


def _run_threads(work, ops):
    """work(k, ops[k]) on one host thread per ctx (ctypes releases the GIL inside libbsk); the first error is re-raised"""
    errs = []

    def guarded(k, op):
        try:
            work(k, op)
        except BaseException as e:  # noqa: BLE001
            errs.append(e)

    th = [threading.Thread(target=guarded, args=(k, op)) for k, op in enumerate(ops)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    if errs:
        raise errs[0]


if __name__ == "__main__":
    sys.exit(main())

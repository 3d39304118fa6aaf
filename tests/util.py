"""Helpers shared by the parity tests: run an operator through the C ABI and through the oracle."""
import oracle
from bigseqkit_b200.api import BskError, Operator

ORACLE_FN = {
    "SeqTransform": oracle.seq, "SubseqTransform": oracle.subseq, "Translate": oracle.translate,
    "Locate": oracle.locate, "Grep": oracle.grep, "Fq2Fa": oracle.fq2fa,
}


def run_lib(lib, op, data, opts, **kw):
    with Operator(op, opts, lib=lib) as o:
        r = o.call(data, **kw)
        return r.data, r.elem_off


def run_oracle(op, data, opts):
    if op == "RmDup":
        d, offs, _ = oracle.rmdup(data, opts)
        return d, offs
    return ORACLE_FN[op](data, opts)


def check_parity(lib, op, data, opts):
    """Bit-exact comparison of output bytes and element offsets, or of the error text."""
    try:
        exp = run_oracle(op, data, opts)
        exp_err = None
    except oracle.OracleError as e:
        exp, exp_err = None, str(e)
    try:
        got = run_lib(lib, op, data, opts)
        got_err = None
    except BskError as e:
        got, got_err = None, str(e)
    assert got_err == exp_err, "error mismatch for %s %r: lib=%r oracle=%r" % (op, opts, got_err, exp_err)
    if exp is not None:
        assert got[0] == exp[0], "output bytes differ for %s %r on %r" % (op, opts, data[:200])
        assert list(got[1]) == list(exp[1]), "element offsets differ for %s %r" % (op, opts)
    return got

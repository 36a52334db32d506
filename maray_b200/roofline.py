"""Algorithmic FP64-pipe work of a scene (the roofline numerator), SURVEY.md section 8(d).

ops(scene) = sum over the non-constant values of the un-hoisted, hash-consed program of w(op), where
w(op) is the number of FP64-pipe SASS instructions (DADD/DMUL/DFMA/DSETP) on the fast path of that
operation, read from the committed SASS of each operation compiled exactly as the back ends compile
it (profiles/sass_ops_r01.txt; regenerate with profiles/dump_op_sass.sh):

    add, mul   1   (DADD / DMUL; --fmad=false, so never a fused pair)
    neg, abs   0   (operand modifiers / integer-pipe sign ops)
    step       1   (DSETP.GE + FSEL)
    min, max   2   (DSETP.GT|LT + DSETP.NAN, 2-4 FSEL on the integer pipe)
    recip      5   (MUFU.RCP64H on the XU pipe + 5 DFMA Newton steps)
    sqrt       8   (MUFU.RSQ64H + 3 DMUL + 5 DFMA)
    sin       15   (1 DSETP + 2 DMUL + 11 DFMA + 1 DADD; F2I/I2F and 3 LDG.128 of coefficients not counted)
    exp       15   (14 DFMA + 1 DADD)
    ln        28   (9 DADD + 15 DFMA + 4 DMUL; MUFU.RCP64H not counted)
    tex        2   (the two `< 0.0` compares; conversions and the byte load are not FP64-pipe work)

The unit is "lane-operations": one FP64-pipe instruction executed for one pixel.  The peak it is
compared with is measured on the same GPU by an FP64 issue-rate microbenchmark
(maray_cuda_fp64_peak: independent DADD/DMUL chains, no FMA), in the same unit.
"""
from __future__ import annotations

OP_WEIGHTS = {
    "n_add": 1, "n_mul": 1, "n_neg": 0, "n_abs": 0, "n_step": 1, "n_min": 2, "n_max": 2,
    "n_recip": 5, "n_sqrt": 8, "n_sin": 15, "n_exp": 15, "n_ln": 28, "n_tex": 2,
}


def fp64_ops_per_pixel(stats: dict) -> int:
    """stats: the dict form of maray_cuda_stats after compile."""
    return sum(w * int(stats[k]) for k, w in OP_WEIGHTS.items())

"""Shared parity checks of the GPU tests (test infrastructure)."""
import numpy as np

# fp32 noise bound for an argmax flip: the 3xTF32 dense products carry ~1e-6 relative error per layer, so two class
# probabilities closer than this (relative, in the float64 oracle) are a numerical tie no fp32 implementation resolves
TIE_REL = 1e-5


def assert_argmax_parity(preds, probs, ref_probs64):
    """argmax parity over ALL rows (north_star: "argmax bit-exact").  Every row must agree with the float64 oracle's
    argmax unless that row is a numerical tie: the oracle's top-2 gap is below TIE_REL relative AND below twice the
    row's own measured |GPU - oracle| error.  Returns (number of mismatching rows, their largest relative gap)."""
    ref_probs64 = np.asarray(ref_probs64, dtype=np.float64)
    ref_pred = ref_probs64.argmax(1)
    mism = np.nonzero(np.asarray(preds) != ref_pred)[0]
    if len(mism) == 0:
        return 0, 0.0
    top = ref_probs64[mism, ref_pred[mism]]
    got = ref_probs64[mism, np.asarray(preds)[mism]]
    gap = top - got
    err = np.abs(np.asarray(probs, dtype=np.float64)[mism] - ref_probs64[mism]).max(axis=1)
    assert (gap <= TIE_REL * top).all(), "argmax mismatch on a row that is not a tie: rel gaps %s" % (gap / top)
    assert (gap <= 2.0 * err + 1e-30).all(), "argmax mismatch larger than the row's numerical error"
    return len(mism), float((gap / top).max())

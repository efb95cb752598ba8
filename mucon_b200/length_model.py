"""Length models with the reference's interface (reference src/core/viterbi/length_model.py).

``PoissonModel`` keeps the reference constructor and ``score`` / ``max_length`` / ``n_classes``
but never runs the 2000-iteration Python loop of length_model.py:65-71 (62 ms per video): it
stores the three per-class numbers the table is made of -- ln m, m, norms -- and the CUDA DP
kernel rebuilds exactly the rows it needs with IEEE mul/sub in the reference's order.  ln() is
NumPy's, i.e. the very function the reference calls, so rows are bit-identical to the
reference's table on the same machine.
"""
import numpy as np

_LOGFACT = {}
_LOGTAIL = np.zeros(2, dtype=np.float64)


def log_factorial_prefix(n):
    """lf[i] = sum_{k=1..i} ln k accumulated sequentially in float64 (length_model.py:67-69)."""
    lf = _LOGFACT.get(n)
    if lf is None:
        lf = np.zeros(n + 1, dtype=np.float64)
        if n >= 1:
            lf[1:] = np.cumsum(np.log(np.arange(1, n + 1, dtype=np.float64)))
        _LOGFACT[n] = lf
    return lf


def _log_tail(top):
    """tail[i] = sum_{k=2..i} ln k, sequential from k = 2 (length_model.py:59-62)."""
    global _LOGTAIL
    if top >= _LOGTAIL.shape[0]:
        n = max(top + 1, 2 * _LOGTAIL.shape[0])
        t = np.zeros(n, dtype=np.float64)
        t[2:] = np.cumsum(np.log(np.arange(2, n, dtype=np.float64)))
        _LOGTAIL = t
    return _LOGTAIL


def poisson_params(mean_lengths, renormalize=True):
    """[C, 3] float64: ln m, m, norms (length_model.py:54-63)."""
    m = np.asarray(mean_lengths, dtype=np.float64)
    out = np.empty(m.shape + (3,), dtype=np.float64)
    if renormalize and m.ndim == 1 and m.size and m.min() > 0.5:
        # every mean rounds to >= 1: no log(0) / 0 * inf to silence -- the same operations in the same order as below,
        # without the error-state context, the clamp and the wrappers (this runs once per video in the evaluator's
        # call pattern: 17 -> 9 us)
        np.log(m, out=out[:, 0])
        out[:, 1] = m
        r = np.rint(m)
        mi = m.astype(np.int64)
        tail = _log_tail(int(mi.max()))
        t = np.log(r)
        t *= r
        t -= r
        t -= tail[mi]
        out[:, 2] = t
        return out
    with np.errstate(divide="ignore", invalid="ignore"):
        out[..., 0] = np.log(m)
        out[..., 1] = m
        if renormalize:
            r = np.round(m)
            mi = np.maximum(m.astype(np.int64), 0)
            tail = _log_tail(int(mi.max(initial=1)))
            out[..., 2] = (r * np.log(r) - r) - tail[mi]
        else:
            out[..., 2] = 0.0
    return out


class LengthModel(object):
    def n_classes(self):
        return 0

    def score(self, length, label):
        return 0.0

    def max_length(self):
        return np.inf


class PoissonModel(LengthModel):
    """Drop-in for core.viterbi.length_model.PoissonModel (length_model.py:42-83)."""

    def __init__(self, model, max_length=2000, renormalize=True):
        super().__init__()
        self.mean_lengths = np.loadtxt(model) if isinstance(model, str) else model
        self.num_classes = self.mean_lengths.shape[0]
        self.max_len = max_length
        self.renormalize = renormalize
        self.params = poisson_params(self.mean_lengths, renormalize)
        self.norms = self.params[:, 2]
        self._table = None
        # The parameter triple (ln m, m, norms) reproduces the reference's table bit for bit when the mean lengths
        # are float64 (what the evaluator passes, evaluators.py:155-165).  For any other dtype the reference's
        # arithmetic runs partly in that dtype (length_model.py:54-71: the norms and l * log(m) - m in the input
        # dtype, the rest promoted by NumPy's rules): then `exact_params` is False and the decoders take explicit
        # rows from `rows_for`, which executes the reference's expressions with the reference's dtypes.
        self.exact_params = np.asarray(self.mean_lengths).dtype == np.float64

    def _low_precision_terms(self):
        """(ln m, m, norms) as the reference computes them for non-float64 means (length_model.py:54-63)."""
        m = np.asarray(self.mean_lengths)
        with np.errstate(divide="ignore", invalid="ignore"):
            norms = np.zeros(m.shape)
            if self.renormalize:
                norms = np.round(m) * np.log(np.round(m)) - np.round(m)             # input dtype
                tail = _log_tail(int(max(int(np.max(m, initial=1)), 1)))
                mi = np.maximum(m.astype(np.int64), 0)
                for c in range(len(m)):                                               # norms[c] = norms[c] - logFak
                    norms[c] = norms[c] - np.float64(tail[mi[c]])
            return np.log(m), m, norms

    def rows_for(self, transcript, fs, J):
        """[N, J] float64 rows[n, j-1] = table[j * fs, transcript[n]] with the reference's own operations and dtypes
        (length_model.py:65-71: l * np.log(m) - m - logFak - norms, logFak an np.float64 scalar)."""
        lnm, m, norms = self._low_precision_terms()
        lf = log_factorial_prefix(self.max_len - 1)
        tr = np.asarray(transcript, dtype=np.int64)
        rows = np.full((tr.shape[0], J), -np.inf, dtype=np.float64)
        with np.errstate(divide="ignore", invalid="ignore"):
            for j in range(1, J + 1):
                l = j * fs
                if l < self.max_len:
                    rows[:, j - 1] = (l * lnm - m - np.float64(lf[l]) - norms)[tr]
        return rows

    @property
    def poisson(self):
        """The reference's full [max_len, C] table, built on demand (vectorised)."""
        if self._table is None and not self.exact_params:
            lnm, m, norms = self._low_precision_terms()
            lf = log_factorial_prefix(self.max_len - 1)
            t = np.zeros((self.max_len, self.num_classes))
            with np.errstate(divide="ignore", invalid="ignore"):
                for l in range(1, self.max_len):
                    t[l, :] = l * lnm - m - np.float64(lf[l]) - norms
            t[0, :] = -np.inf
            self._table = t
        if self._table is None:
            lf = log_factorial_prefix(self.max_len - 1)
            L = np.arange(self.max_len, dtype=np.float64)[:, None]
            p = self.params
            with np.errstate(invalid="ignore"):
                t = ((L * p[None, :, 0] - p[None, :, 1]) - lf[:, None]) - p[None, :, 2]
            t[0, :] = -np.inf
            self._table = t
        return self._table

    def n_classes(self):
        return self.num_classes

    def score(self, length, label):
        if length >= self.max_len or length <= 0:
            return -np.inf
        if not self.exact_params:
            return self.poisson[length, label]
        lm, m, nrm = self.params[label]
        return ((length * lm - m) - log_factorial_prefix(self.max_len - 1)[length]) - nrm

    def max_length(self):
        return self.max_len

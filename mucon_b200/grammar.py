"""Grammars with the reference's interface (reference src/core/viterbi/grammar.py).

The CUDA decoder only needs the list of candidate transcripts a grammar admits; these classes
keep that list (``candidates``) next to the reference's ``successors`` prefix tree so that code
written against the reference protocol (n_classes / possible_successors / score) still works.
"""
import numpy as np


class Grammar(object):
    def score(self, context, label):
        return 0.0

    def n_classes(self):
        return 0

    def start_symbol(self):
        return -1

    def end_symbol(self):
        return -2

    def possible_successors(self, context):
        return set()

    def update_context(self, context, label):
        return context + (label,)


class _TranscriptSetGrammar(Grammar):
    def _install(self, transcripts, num_classes):
        self.num_classes = num_classes
        self.candidates = []
        seen = set()
        self._all = []          # every transcript in the order given (duplicates included): what builds the prefix tree
        self._succ = None
        for tr in transcripts:
            tr = [int(x) for x in tr]
            self._all.append(tr)
            if tuple(tr) not in seen:
                seen.add(tuple(tr))
                self.candidates.append(tr)

    @property
    def successors(self):
        """The reference's prefix tree (context tuple -> set of next labels), built on first use: the CUDA decoder reads
        `candidates`, and the evaluator constructs a new grammar for every video (evaluators.py:167)."""
        if self._succ is None:
            succ = {}
            for tr in self._all:
                path = tr + [self.end_symbol()]
                for i, nxt in enumerate(path):
                    # built exactly like grammar.py:150-154 ({x}.union(old)): the ITERATION order of these sets decides
                    # which of two candidates with equal scores the reference returns (tie_ranks below)
                    ctx = (self.start_symbol(),) + tuple(path[:i])
                    succ[ctx] = {nxt}.union(succ.get(ctx, set()))
            self._succ = succ
        return self._succ

    @successors.setter
    def successors(self, value):
        self._succ = value

    def n_classes(self):
        return self.num_classes

    def possible_successors(self, context):
        return self.successors.get(context, set())

    def score(self, context, label):
        return 0.0 if label in self.possible_successors(context) else -np.inf


class SingleTranscriptGrammar(_TranscriptSetGrammar):
    """grammar.py:196-217 -- exactly one admissible transcript."""

    def __init__(self, transcript, n_classes):
        self._install([transcript], n_classes)


class ModifiedPathGrammar(_TranscriptSetGrammar):
    """grammar.py:178-191 -- every transcript of a given list."""

    def __init__(self, transcripts, num_classes):
        self._install(transcripts, num_classes)


class PathGrammar(_TranscriptSetGrammar):
    """grammar.py:143-175 -- transcripts read from a text file, one per line."""

    def __init__(self, transcript_file, label2index_map):
        with open(transcript_file, "r") as f:
            lines = f.read().split("\n")[0:-1]
        self._install([[label2index_map[w] for w in line.split()] for line in lines], len(label2index_map))


def lower_grammar(grammar):
    """Candidate transcripts of any grammar that follows the reference protocol with 0/-inf
    scores and a finite prefix tree (ours, or the reference's Single/Path grammars)."""
    cands = getattr(grammar, "candidates", None)
    if cands is not None:
        return [list(c) for c in cands]
    succ = getattr(grammar, "successors", None)
    if not isinstance(succ, dict):
        raise TypeError(
            f"{type(grammar).__name__} cannot be lowered to candidate transcripts; the CUDA decoder "
            "supports SingleTranscriptGrammar / PathGrammar / ModifiedPathGrammar (there is no CPU fallback)")
    start, end = grammar.start_symbol(), grammar.end_symbol()
    out = []
    stack = [((start,), [])]
    while stack:
        ctx, tr = stack.pop()
        for nxt in sorted(succ.get(ctx, ()), reverse=True):
            if nxt == end:
                out.append(tr)
            else:
                stack.append((ctx + (nxt,), tr + [int(nxt)]))
    if not out:
        raise ValueError("grammar admits no transcript")
    return out


def tie_ranks(successors, candidates, start=-1):
    """[len(candidates), 2] int32 for mucon_viterbi_select_ranked: the place of each candidate's final hypothesis in
    the reference's insertion-ordered hypothesis dict, as far as the grammar decides it (viterbi.py:93-138: successors
    are visited in the iteration order of the grammar's sets, `>=` lets the later hypothesis win an exact tie).
    Column 0: place of the first label among the start context's successors; column 1: dense rank of the places of the
    remaining labels, compared as sequences (a proper prefix is earlier).  The data-dependent middle key -- last
    segment length + transcript length -- is added on the device."""
    place = {}
    get = place.get
    seqs = []
    for tr in candidates:
        ctx = (start,)
        r = []
        for x in tr:
            x = int(x)
            pl = get(ctx)
            if pl is None:
                pl = place[ctx] = {int(y): i for i, y in enumerate(successors.get(ctx, ()))}
            r.append(pl[x])
            ctx = ctx + (x,)
        seqs.append(tuple(r))
    rest = {t: i for i, t in enumerate(sorted({q[1:] for q in seqs}))}
    out = np.zeros((len(seqs), 2), dtype=np.int32)
    for i, q in enumerate(seqs):
        out[i, 0] = q[0] if q else 0
        out[i, 1] = rest[q[1:]]
    return out


def tie_ranks_for_lists(candidates, start=-1, end=-2):
    """tie_ranks for a plain list of transcripts = what ModifiedPathGrammar(candidates) (grammar.py:178-191) implies."""
    succ = {}
    empty = frozenset()
    get = succ.get
    for tr in candidates:
        ctx = (start,)
        for nxt in tr:
            nxt = int(nxt)
            succ[ctx] = {nxt}.union(get(ctx, empty))     # the reference's statement (grammar.py:185-189): order matters
            ctx = ctx + (nxt,)
        succ[ctx] = {end}.union(get(ctx, empty))
    return tie_ranks(succ, candidates, start)

"""GPU scratch tool: where the c1 drop-in decode's wall time goes."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mucon_b200 import PoissonModel, SingleTranscriptGrammar
from mucon_b200.viterbi import Viterbi
from tests import synth
dev = torch.device("cuda:0")
rng = np.random.default_rng(0)
tr = [0, 5, 7, 5, 12, 0]
lp, _ = synth.planted_logp(rng, 2000, 48, tr, np.float32)
means = synth.class_means(rng.dirichlet(5 * np.ones(6)).astype(np.float32), tr, 48, 2000)
dec = Viterbi(SingleTranscriptGrammar(tr, 48), PoissonModel(means), frame_sampling=30, device=dev)
for _ in range(10): dec.decode(lp)
def t(fn, n=200):
    t0 = time.perf_counter()
    for _ in range(n): fn()
    return (time.perf_counter() - t0) / n * 1e6
print("decode only (objects reused)        %.1f us" % t(lambda: dec.decode(lp)))
def full():
    dec.grammar = SingleTranscriptGrammar(tr, 48); dec.length_model = PoissonModel(means); dec.decode(lp)
print("grammar + PoissonModel + decode     %.1f us" % t(full))
eng = dec._eng(); lm = dec.length_model
print("_decode_single alone                %.1f us" % t(lambda: dec._decode_single(eng, lp, tr, lm, 30, 2000)))
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for _ in range(200): full()
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(8)

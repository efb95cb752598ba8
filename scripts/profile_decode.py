"""GPU scratch tool: where the host time of one drop-in Viterbi.decode call goes (cProfile)."""
import cProfile, os, pstats, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mucon_b200 import PoissonModel, SingleTranscriptGrammar
from mucon_b200.viterbi import Viterbi
from tests import synth
rng = np.random.default_rng(0)
tr = [0, 5, 7, 5, 12, 0]
lp, _ = synth.planted_logp(rng, 2000, 48, tr, np.float32)
means = synth.class_means(rng.dirichlet(5 * np.ones(6)).astype(np.float32), tr, 48, 2000)
dec = Viterbi(SingleTranscriptGrammar(tr, 48), PoissonModel(means), frame_sampling=30)
for _ in range(20):
    dec.decode(lp)
pr = cProfile.Profile()
pr.enable()
for _ in range(200):
    dec.grammar = SingleTranscriptGrammar(tr, 48)
    dec.length_model = PoissonModel(means)
    dec.decode(lp)
pr.disable()
st = pstats.Stats(pr)
st.sort_stats("cumulative").print_stats(28)

"""Mints tests/golden/eval_lengths.npz: the class-mean-length block of the reference evaluator
(/root/reference/src/mucon/evaluators.py, the statements from `actions = one_hot(` to
`lengths[lengths == 0] = 1`, plus its `one_hot` helper) executed UNMODIFIED -- the statements are
read from the reference file at mint time and exec'd on seeded inputs; nothing is copied into the
repo.  Re-run with:  python tests/golden/make_golden_eval.py
"""
import os
import textwrap

import numpy as np
import torch

REF = "/root/reference/src/mucon/evaluators.py"
HERE = os.path.dirname(os.path.abspath(__file__))


def reference_block():
    src = open(REF).read().split("\n")
    a = next(i for i, l in enumerate(src) if l.strip().startswith("def one_hot("))
    helper = "\n".join(src[a:a + 2])
    b = next(i for i, l in enumerate(src) if l.strip().startswith("actions = one_hot("))
    e = next(i for i, l in enumerate(src) if i > b and l.strip().startswith("lengths[lengths == 0] = 1"))
    return helper, textwrap.dedent("\n".join(src[b:e + 1]))


def main():
    helper, block = reference_block()
    rng = np.random.default_rng(7)
    cases = {}
    for ci in range(8):
        C = int(rng.integers(4, 49))
        N = int(rng.integers(1, 13))
        tr = rng.integers(0, C, N)
        if ci == 1:
            tr = np.array([0, 5, 7, 5, 12, 0]) % C
        rel = torch.softmax(torch.from_numpy(rng.normal(size=len(tr)).astype(np.float32)) * 2, 0)
        if ci == 2:
            rel = rel.clone(); rel[0] = 0.0  # an exact zero: replaced by 1
        T = int(rng.integers(60, 9000))
        ns = {"np": np, "tensor_to_numpy": lambda t: t.detach().cpu().numpy(),
              "predicted_transcript_s_head_list": tr.tolist(), "number_of_action_classes": C,
              "predicted_relative_lengths": rel, "feature_length": T}
        exec(helper, ns)
        exec(block, ns)
        cases[f"tr{ci}"] = tr.astype(np.int64)
        cases[f"rel{ci}"] = rel.numpy()
        cases[f"T{ci}"] = np.int64(T)
        cases[f"C{ci}"] = np.int64(C)
        cases[f"lengths{ci}"] = ns["lengths"]
    np.savez_compressed(os.path.join(HERE, "eval_lengths.npz"), n=np.int64(8), **cases)
    print("wrote eval_lengths.npz;", block.split("\n")[0], "...", block.split("\n")[-1])


if __name__ == "__main__":
    main()

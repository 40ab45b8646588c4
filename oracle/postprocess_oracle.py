"""CPU oracle for the validation metrics of cadia-lvl/ss_asr.

TEST INFRASTRUCTURE ONLY.  Nothing in `ss_asr_b200/` may import this module; only `tests/` does, as the checker.

Pure-Python / numpy restatement, written from the reference's semantics:

  * calc_acc ................. /root/reference/src/postprocess.py:7-29
  * calc_err ................. /root/reference/src/postprocess.py:31-50
  * trim_eos ................. /root/reference/src/postprocess.py:68-75
  * Mapper.translate ......... /root/reference/src/ASRDataset.py:228-252
  * ASRTrainer.valid body .... /root/reference/src/trainer.py:472-494

The edit distance lives in the third-party package `editdistance` (requirements.txt; not vendored, not installed here):
`editdistance.eval(a, b)` is the unit-cost Levenshtein distance between two sequences of hashables, restated in
`levenshtein` below.

Parity pinning: the reference ships no vectors for this path.  The oracle is pinned against the UNMODIFIED
`postprocess.calc_acc` / `calc_err` and `ASRDataset.Mapper` executed through oracle/ref_shim.py (whose `editdistance`
stub is the same published algorithm) on seeded cases that include ties, early EOS, empty words, repeated spaces and
predictions shorter / longer than the label: tests/golden/postprocess.npz (made by tests/golden/make_golden_postprocess.py),
re-checked by tests/test_oracle_golden.py.
"""
import numpy as np

TOKENS = '<>$' + 'abcdefghijklmnoprstuvxy0123456789' + 'áéíóúýæöþð' + ' .,?'   # preprocess.py:17-27
SOS_TKN, EOS_TKN = '<', '>'


def levenshtein(a, b):
    a, b = list(a), list(b)
    prev = list(range(len(b) + 1))
    for i in range(1, len(a) + 1):
        cur = [i] + [0] * len(b)
        for j in range(1, len(b) + 1):
            cur[j] = min(prev[j] + 1, cur[j - 1] + 1, prev[j - 1] + (0 if a[i - 1] == b[j - 1] else 1))
        prev = cur
    return prev[len(b)]


def trim_eos(seq):
    out = []
    for c in seq:
        out.append(int(c))
        if int(c) == 1:
            break
    return out


def translate(seq, tokens=TOKENS):
    s = ''.join(tokens[c] for c in trim_eos(seq))
    return s.replace(SOS_TKN, '').replace(EOS_TKN, '')


def argmax_tokens(predict):
    return np.argmax(np.asarray(predict), axis=-1)


def utterance_stats(predict, label, tokens=TOKENS):
    """-> int array [B, 4]: correct, total (calc_acc); word edit distance, label words (calc_err)."""
    pred = argmax_tokens(predict)
    label = np.asarray(label)
    out = np.zeros((pred.shape[0], 4), dtype=np.int64)
    for b in range(pred.shape[0]):
        correct = total = 0
        for pp, ll in zip(pred[b], label[b]):
            if ll == 0:
                break
            correct += int(pp == ll)
            total += 1
        pw = translate(pred[b], tokens).split(' ')
        lw = translate(label[b], tokens).split(' ')
        out[b] = (correct, total, levenshtein(pw, lw), len(lw))
    return out


def calc_acc(predict, label):
    st = utterance_stats(predict, label)
    accs = [float(c) / t for c, t, _, _ in st.tolist()]
    return sum(accs) / len(accs)


def calc_err(predict, label, tokens=TOKENS):
    st = utterance_stats(predict, label, tokens)
    ds = [float(d) / n for _, _, d, n in st.tolist()]
    return sum(ds) / len(ds)


def synth_cases(seed=0, B=24, U=37, L=17, C=50):
    """Seeded prediction / label batch that exercises the edge cases: frequent spaces (empty words, repeated words),
    early EOS in predictions and labels, exact ties in the prediction rows, predictions equal to the label, labels
    that start with padding's neighbour tokens, U > L and (with swapped arguments) U < L."""
    rng = np.random.RandomState(seed)
    label = np.zeros((B, L), dtype=np.int64)
    pred_tok = np.zeros((B, U), dtype=np.int64)
    alphabet = np.array([3, 4, 5, 46, 46, 46, 7, 2, 47, 49], dtype=np.int64)
    for b in range(B):
        n = int(rng.randint(1, L)) if L > 1 else 1       # label tokens before EOS
        label[b, :n] = alphabet[rng.randint(0, len(alphabet), n)]
        if n < L:
            label[b, n] = 1
        mode = b % 4
        p = alphabet[rng.randint(0, len(alphabet), U)]
        if mode == 0:                                   # mostly right, some substitutions / one deletion
            m = min(n, U - 1)
            p[:m] = label[b, :m]
            p[m] = 1
            flip = rng.rand(m) < 0.2
            p[:m][flip] = alphabet[rng.randint(0, len(alphabet), int(flip.sum()))]
        elif mode == 1:                                 # no EOS at all: the whole U steps are the hypothesis
            pass
        elif mode == 2:                                 # EOS straight away: empty hypothesis
            p[0] = 1
        else:                                           # SOS / EOS-char ids inside the hypothesis, EOS late
            p[rng.randint(0, U, 3)] = 0
            p[max(U - 2, 0)] = 1
        pred_tok[b] = p
    predict = rng.randn(B, U, C).astype(np.float32)
    for b in range(B):
        for u in range(U):
            predict[b, u, pred_tok[b, u]] = 9.0
            if (b + u) % 5 == 0 and pred_tok[b, u] + 1 < C:     # exact tie: the first maximum must win
                predict[b, u, pred_tok[b, u] + 1] = 9.0
    return predict, label, pred_tok

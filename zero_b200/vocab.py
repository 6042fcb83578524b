"""Vocabulary of the data path (reference vocab.py:8-84): ids 0 / 1 / 2 are <pad> / <unk> / <eos> (vocab.py:20-22),
`to_id` appends <eos> (vocab.py:69-73), unknown tokens map to <unk>, unknown ids print as <unk>.

The id conventions are part of the hot-path contract: masks are `id != 0` (models/transformer.py:16) and beam
search stops on id 2 (search.py:192).
"""
from __future__ import annotations

PAD, UNK, EOS = "<pad>", "<unk>", "<eos>"


class Vocab(object):
    def __init__(self, vocab_file=None, tokens=None):
        self._tok2id = {}
        self._id2tok = []
        self._count = {}
        for t in (PAD, UNK, EOS):
            self.insert(t)
        if vocab_file is not None:
            self.load_vocab(vocab_file)
        if tokens is not None:
            for t in tokens:
                self.insert(t)

    # -- construction ------------------------------------------------------------------------------
    def insert(self, token):
        if token not in self._tok2id:
            self._tok2id[token] = len(self._id2tok)
            self._id2tok.append(token)
            self._count[token] = 0
        self._count[token] += 1

    def sort_vocab(self):
        """Most frequent first behind the three specials (vocab.py:54-62; Python's sort is stable, so equally
        frequent tokens keep their order of first appearance, as in the reference)."""
        ranked = sorted(self._count.items(), key=lambda x: -x[1])
        self._tok2id, self._id2tok = {}, []
        for t in (PAD, UNK, EOS):
            self.insert(t)
        for tok, _ in ranked:
            self.insert(tok)

    def load_vocab(self, vocab_file):
        with open(vocab_file, "r") as f:
            for line in f:
                self.insert(line.strip())

    def save_vocab(self, vocab_file, size=10 ** 6):
        with open(vocab_file, "w") as f:
            for tok in self._id2tok[:int(size)]:
                f.write(tok + "\n")

    # -- lookups -----------------------------------------------------------------------------------
    def size(self):
        return len(self._id2tok)

    def get_id(self, token):
        return self._tok2id.get(token, 1)

    def get_token(self, idx):
        idx = int(idx)
        return self._id2tok[idx] if 0 <= idx < len(self._id2tok) else UNK

    def to_id(self, tokens, append_eos=True):
        ids = [self.get_id(t) for t in tokens]
        if append_eos:
            ids.append(2)
        return ids

    def to_tokens(self, ids):
        return [self.get_token(i) for i in ids]

    @staticmethod
    def pad():
        return 0

    @staticmethod
    def unk():
        return 1

    @staticmethod
    def eos():
        return 2


def main(argv=None):
    """`python -m zero_b200.vocab [--size N] corpus vocab_file`: the command-line form of the reference's vocabulary
    builder (vocab.py:86-103) — count the whitespace tokens of `corpus`, order them by frequency behind the three
    specials, write at most N entries, one per line."""
    import argparse
    cli = argparse.ArgumentParser(prog="zero_b200.vocab", description="build a vocabulary file from a tokenised corpus")
    cli.add_argument("--size", type=int, default=10 ** 6, help="keep at most this many entries (specials included)")
    cli.add_argument("input", help="tokenised text, one sentence per line")
    cli.add_argument("output", help="vocabulary file to write")
    opt = cli.parse_args(argv)
    vocab = Vocab()
    with open(opt.input, "r") as corpus:
        for sentence in corpus:
            for word in sentence.split():
                vocab.insert(word)
    vocab.sort_vocab()
    vocab.save_vocab(opt.output, opt.size)
    print("%d distinct tokens (specials included) in %s -> %s" % (vocab.size(), opt.input, opt.output))
    return vocab


if __name__ == "__main__":
    main()

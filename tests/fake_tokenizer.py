"""A toy sentencepiece-style HuggingFace tokenizer: the methods GenerationConfig.convert_stop_bad_words_to_ids and the
reference's Tokenizer.indexes_containing_token use (get_vocab / encode / decode / convert_ids_to_tokens).  Shared by the
CPU tests and tools/make_stopwords_golden.py, which drives the REFERENCE's own methods over the same vocabulary."""


class FakeTokenizer:
    VOCAB = {"<s>": 0, "</s>": 1, "▁Hello": 2, "▁": 3, "!": 4, "!!": 5, "▁wor": 6, "ld": 7, "\n": 8, "\n\n": 9,
             "a!": 10, "b!": 11, "c!": 12, "d!": 13, "<|im_end|>": 14, "▁<|im_end|>": 15, "Ġx": 16}

    def __init__(self, vocab=None):
        self.vocab = dict(vocab if vocab is not None else self.VOCAB)
        self.inv = {i: t for t, i in self.vocab.items()}

    def get_vocab(self):
        return dict(self.vocab)

    def __len__(self):
        return len(self.vocab)

    def convert_ids_to_tokens(self, ids):
        return self.inv[ids] if isinstance(ids, int) else [self.inv[i] for i in ids]

    def encode(self, text, add_special_tokens=True):
        order = sorted(self.vocab, key=len, reverse=True)
        text = text.replace(" ", "▁")
        out = []
        while text:
            for tok in order:
                if text.startswith(tok):
                    out.append(self.vocab[tok])
                    text = text[len(tok):]
                    break
            else:                       # a character outside the vocabulary: an id beyond it (byte-fallback stand-in)
                out.append(1000 + ord(text[0]))
                text = text[1:]
        return out

    def decode(self, ids, **kw):
        ids = [ids] if isinstance(ids, int) else ids
        return "".join(self.inv[i] if i in self.inv else chr(i - 1000) for i in ids).replace("▁", " ")

"""EDA text augmentation for DeCLIP's second caption view — prototype/model/declip.py:154-155,203-212 picks one of
`EDA().synonym_replacement / random_swap / random_deletion` per caption.

`EDA` there is `textaugment.EDA`, a third-party dependency that is not part of the reference tree (it is imported at
declip.py:16 and not pinned in requirements.txt); its three operations are the published "Easy Data Augmentation"
procedures of Wei & Zou (EMNLP 2019), restated here:
  * random_swap(sentence, n=1)      : n times, swap the words at two distinct random positions (gives up after 3 tries)
  * random_deletion(sentence, p=0.1): drop each word with probability p; a one-word sentence is returned as is and an
                                      emptied sentence becomes one random word of the original
  * synonym_replacement(sentence, n=1): pick up to n distinct non-stop-words that have synonyms and replace every
                                      occurrence with a random synonym.  textaugment queries NLTK WordNet, a corpus
                                      download that is not available offline: the synonym source is pluggable
                                      (`synonyms=` mapping word -> list of words, or a callable); without one the
                                      sentence is returned unchanged, exactly what WordNet yields for a word it lacks.
All randomness comes from one `random.Random(seed)` so a run is reproducible (textaugment uses the global `random`)."""
import random

# the stop-word list of the EDA paper's reference implementation (function words are never replaced)
STOP_WORDS = frozenset(
    "i me my myself we our ours ourselves you your yours yourself yourselves he him his himself she her hers herself it "
    "its itself they them their theirs themselves what which who whom this that these those am is are was were be been "
    "being have has had having do does did doing a an the and but if or because as until while of at by for with about "
    "against between into through during before after above below to from up down in out on off over under again further "
    "then once here there when where why how all any both each few more most other some such no nor not only own same so "
    "than too very s t can will just don should now".split())


class EDA:
    def __init__(self, synonyms=None, stop_words=STOP_WORDS, random_state=None):
        self.stop_words = stop_words
        self.rng = random.Random(random_state)
        if synonyms is None:
            self._syn = lambda w: ()
        elif callable(synonyms):
            self._syn = synonyms
        else:
            self._syn = lambda w: synonyms.get(w, ())

    # ---------------------------------------------------------------- the three operations
    def synonym_replacement(self, sentence, n=1):
        words = sentence.split()
        candidates = list(dict.fromkeys(w for w in words if w.lower() not in self.stop_words))
        self.rng.shuffle(candidates)
        replaced = 0
        for w in candidates:
            syn = [s for s in self._syn(w.lower()) if s != w.lower()]
            if syn:
                pick = self.rng.choice(syn)
                words = [pick if x == w else x for x in words]
                replaced += 1
            if replaced >= n:
                break
        return " ".join(words)

    def random_swap(self, sentence, n=1):
        words = sentence.split()
        if len(words) < 2:
            return sentence
        for _ in range(n):
            i = self.rng.randrange(len(words))
            j, tries = i, 0
            while j == i:
                j = self.rng.randrange(len(words))
                tries += 1
                if tries > 3:
                    break
            if j != i:
                words[i], words[j] = words[j], words[i]
        return " ".join(words)

    def random_deletion(self, sentence, p=0.1):
        words = sentence.split()
        if len(words) <= 1:
            return sentence
        kept = [w for w in words if self.rng.random() > p]
        if not kept:
            return self.rng.choice(words)
        return " ".join(kept)

    # ---------------------------------------------------------------- declip.py:203-212
    def augment(self, caption):
        op = self.rng.choice((self.synonym_replacement, self.random_swap, self.random_deletion))
        out = op(caption)
        return " ".join(out) if isinstance(out, list) else out

    def augment_batch(self, captions):
        return [self.augment(c) for c in captions]

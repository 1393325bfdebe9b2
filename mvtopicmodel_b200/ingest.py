"""Host-side ingestion for the sampling path: raw text / delimiter-separated side views ->
the `InstanceList`s that `FastQMVWVParallelTopicModel.addInstances` takes.

Follows the reference's `SciTopicFlow.ImportInstancesWithNewPipes` (S:1800-1930) and `GenerateStoplist` (S:631-730) with the
MALLET 2.0.8 pipes it instantiates.  MALLET ships as a binary jar; the rules below were recovered from its bytecode with
tools/jclass.py (each function names the class and method it restates):

    Input2CharSequence -> CharSequenceLowercase -> SimpleTokenizer(stoplists/en.txt) -> StringList2FeatureSequence   (text view)
    CSV2FeatureSequence(delimiter)                                                                                 (side views)

Nothing here touches the device; it is plain Python over strings.
"""
import re
import unicodedata

import numpy as np

from .model import Instance, InstanceList

# java.lang.Character.getType() codes of the Unicode general categories SimpleTokenizer.pipe looks at
_JAVA_TYPE = {"Lu": 1, "Ll": 2, "Lt": 3, "Lm": 4, "Lo": 5, "Mn": 6, "Me": 7, "Mc": 8, "Nd": 9, "Nl": 10, "No": 11, "Zs": 12, "Zl": 13,
              "Zp": 14, "Cc": 15, "Cf": 16, "Co": 18, "Cs": 19, "Pd": 20, "Ps": 21, "Pe": 22, "Pc": 23, "Po": 24, "Sm": 25, "Sc": 26,
              "Sk": 27, "So": 28, "Pi": 29, "Pf": 30, "Cn": 0}
_APPEND = frozenset((1, 2, 3, 4, 5, 6, 7, 8))          # letters (all five kinds) and marks extend the current token
_CLOSE = frozenset((12, 13, 14, 20, 21, 22, 23, 24, 29, 30))   # separators and punctuation close it
_MAX_TOKEN = 1000                                     # `new int[1000]` token buffer: a full buffer is flushed as a token


class Alphabet:
    """cc.mallet.types.Alphabet: insertion-ordered symbol table (lookupIndex adds unseen entries)."""

    def __init__(self):
        self.index, self.entries = {}, []

    def lookup_index(self, s, add=True):
        i = self.index.get(s)
        if i is None:
            if not add:
                return -1
            i = len(self.entries)
            self.index[s] = i
            self.entries.append(s)
        return i

    def lookup_object(self, i):
        return self.entries[i]

    def __len__(self):
        return len(self.entries)

    def __contains__(self, s):
        return s in self.index


def load_stoplist(path):
    """SimpleTokenizer.<init>(File): every line of the UTF-8 file, untrimmed, is one stop word."""
    with open(path, encoding="utf-8", errors="replace") as f:
        return set(line.rstrip("\n").rstrip("\r") for line in f)


def simple_tokenize(text, stoplist=frozenset()):
    """cc.mallet.pipe.SimpleTokenizer.pipe (bytecode pc 0-383).  Per code point: letters of any case, marks and '_' extend the
    token; separators (Zs/Zl/Zp) and punctuation (Pd/Ps/Pe/Pc/Po/Pi/Pf) close it; EVERYTHING ELSE -- digits, symbols, and
    control characters such as TAB and NEWLINE -- is dropped without closing it ("abc1\\tdef" is the single token "abcdef").
    A closed token is kept unless the stop list holds it; a token is also flushed when it fills the 1000-code-point buffer."""
    out, cur = [], []

    def flush():
        if cur:
            tok = "".join(cur)
            if tok not in stoplist:
                out.append(tok)
            cur.clear()

    for ch in text:
        t = _JAVA_TYPE.get(unicodedata.category(ch), 0)
        if t == 1 or t == 2 or ch == "_":
            cur.append(ch)
        elif t in _CLOSE:
            flush()
        elif t in _APPEND:
            cur.append(ch)
        if len(cur) == _MAX_TOKEN:
            flush()
    flush()
    return out


# words GenerateStoplist always stops (S:683-718): fragments left by a lossy PDF-to-text step upstream of the reference
_FRAGMENTS = ("tion ing ment ytem wth whch nfrmatn uer ther frm hypermeda anuae dcument tudent appcatn tructure prram den aed cmputer "
              "mre cence tures ture ments cations tems tem tional ity ware opment guage niques").split()
_WORD_OK = re.compile(r"^(?!.*(-[^-]*-|_[^_]*_))[A-Za-z0-9][\w-]*[A-Za-z0-9]$", re.ASCII)      # S:678 (java \w is ASCII)
_WORD_BAD = ("cid", "italic", "null", "usepackage", "fig")


def generate_stoplist(texts, stoplist, prune_count, doc_proportion_max_cutoff=10.0, preserve_case=False):
    """SciTopicFlow.GenerateStoplist (S:631-730): tokenises every document once with a copy of the tokenizer, then ADDS to
    `stoplist` (in place, as the reference mutates its tokenizer): alphabet entries failing the word-shape test S:678, the
    fixed fragment list, entries seen fewer than `prune_count` times (FeatureCountPipe.addPrunedWordsToStoplist: count <
    prune_count) and -- only when the cut-off is below 1 -- entries whose document frequency exceeds it
    (FeatureDocFreqPipe.addPrunedWordsToStoplist: df / numInstances > cutoff).  Returns the stop list."""
    base = set(stoplist)
    alphabet, counts, dfs = Alphabet(), [], []
    n_docs = 0
    for text in texts:
        n_docs += 1
        toks = simple_tokenize(text if preserve_case else text.lower(), base)
        seen = set()
        for tok in toks:
            i = alphabet.lookup_index(tok)
            if i == len(counts):
                counts.append(0); dfs.append(0)
            counts[i] += 1
            if i not in seen:
                seen.add(i); dfs[i] += 1
    for w in alphabet.entries:
        if not _WORD_OK.match(w) or len(w) < 3 or any(b in w for b in _WORD_BAD):
            stoplist.add(w)
    stoplist.update(_FRAGMENTS)
    if prune_count > 0:
        for i, w in enumerate(alphabet.entries):
            if counts[i] < prune_count:
                stoplist.add(w)
    if doc_proportion_max_cutoff < 1.0:
        for i, w in enumerate(alphabet.entries):
            if dfs[i] / n_docs > doc_proportion_max_cutoff:
                stoplist.add(w)
    return stoplist


def _java_split(s, regex):
    """String.split(regex): trailing empty strings are removed; a leading empty string is kept (for non-empty input)."""
    parts = re.split(regex, s)
    while parts and parts[-1] == "":
        parts.pop()
    return parts


def csv_to_features(data, alphabet, delimiter=",", stoplist=frozenset()):
    """org.madgik.utils.CSV2FeatureSequence.pipe (CSV2FeatureSequence.java:63-98): split on the delimiter regex, keep tokens
    LONGER THAN THREE characters whose lower-case form is not on the stop list (the token itself is stored with its case)."""
    if not data:
        return np.zeros(0, dtype=np.int32)
    ids = [alphabet.lookup_index(tok) for tok in _java_split(data, delimiter) if len(tok) > 3 and tok.lower() not in stoplist]
    return np.asarray(ids, dtype=np.int32)


def _java_round(x):
    """Math.round(double): floor(x + 0.5)."""
    return int(np.floor(x + 0.5))


def import_instances(instance_buffer, num_modalities, prune_cnt_perc=0.002, prune_lbl_cnt_perc=0.002, prune_max_perc=10.0,
                     ignore_text=False, csv_delimiter=",", text_stoplist=frozenset(), csv_stoplist=frozenset(), pubmed=False):
    """SciTopicFlow.ImportInstancesWithNewPipes (S:1800-1930).  instance_buffer[m] is a list of (name, data string) pairs.
    Returns (InstanceList per modality, Alphabet per modality).

    Text view (m = 0 unless ignore_text): stop list = `text_stoplist` grown by generate_stoplist with
    prune_count = round(#docs * prune_cnt_perc); then lower-case -> SimpleTokenizer -> alphabet lookup (S:1809-1845).
    Side views: CSV2FeatureSequence, then -- if prune_lbl_cnt_perc > 0 and the view has more than 10 instances -- features
    whose corpus count is below round(#instances * prune_lbl_cnt_perc) (x4 for modality 3 of the PubMed experiment) are
    dropped and the alphabet is rebuilt in order of first surviving occurrence (FeatureSequence.prune(counts, newAlphabet,
    min): keeps a position iff counts[feature] >= min), S:1859-1905."""
    lists, alphabets = [None] * num_modalities, [None] * num_modalities
    first_side = 0 if ignore_text else 1
    if not ignore_text:
        stop = set(text_stoplist)
        prune_count = _java_round(len(instance_buffer[0]) * prune_cnt_perc)
        generate_stoplist((d for _, d in instance_buffer[0]), stop, prune_count, prune_max_perc, False)
        alpha = Alphabet()
        il = InstanceList(alphabet_size=None)
        for name, data in instance_buffer[0]:
            il.append(Instance(name, [alpha.lookup_index(t) for t in simple_tokenize(data.lower(), stop)]))
        lists[0], alphabets[0] = il, alpha
    for m in range(first_side, num_modalities):
        alpha = Alphabet()
        feats = [(name, csv_to_features(data, alpha, csv_delimiter, csv_stoplist)) for name, data in instance_buffer[m]]
        if prune_lbl_cnt_perc > 0 and len(feats) > 10:
            counts = np.zeros(len(alpha), dtype=np.float64)
            for _, f in feats:
                np.add.at(counts, f, 1.0)                                      # FeatureSequence.addFeatureWeightsTo
            pr = _java_round(len(instance_buffer[m]) * prune_lbl_cnt_perc)
            if m == 3 and pubmed:
                pr *= 4
            new_alpha, pruned = Alphabet(), []
            for name, f in feats:
                keep = f[counts[f] >= pr] if len(f) else f
                pruned.append((name, np.asarray([new_alpha.lookup_index(alpha.lookup_object(int(i))) for i in keep], dtype=np.int32)))
            alpha, feats = new_alpha, pruned
        il = InstanceList(alphabet_size=None)
        for name, f in feats:
            il.append(Instance(name, f))
        lists[m], alphabets[m] = il, alpha
    for m in range(num_modalities):
        if lists[m] is not None:
            lists[m]._alphabet_size = len(alphabets[m])
            lists[m].alphabet = alphabets[m]
    return lists, alphabets


def read_sms_collection(path):
    """SampleData/SMSSpamCollection2.txt: `id TAB label TAB text` per line -> [(id, text)] (BASELINE configs[0])."""
    out = []
    with open(path, encoding="utf-8", errors="replace") as f:
        for line in f:
            parts = line.rstrip("\n").split("\t", 2)
            if len(parts) == 3:
                out.append((parts[0], parts[2]))
    return out

"""Text formats of the reference's model readers, written from / read into plain arrays so
that results can be diffed against files a Java run of hmetaxa/MVTopicModel left on disk:

    printState            M:3269-3320   gzip text, one line per token
    printTypeTopicCounts  M:2076-2102   `type word t:count ...` per word type
    printTopicWordWeights M:2104-2129   `topic \\t word \\t beta+count`
    getSortedWords / getTopWords / displayTopWords   M:1792-1890 (MALLET IDSorter order)

Java's number -> text rules are restated here: Double.toString (JLS / JDK >= 19 shortest-digits form) and
NumberFormat.getInstance() with at most five fraction digits (M:221-222).  Pure host code; works on numpy arrays.
"""
import gzip
import io
import math
from decimal import ROUND_HALF_EVEN, Context, Decimal

import numpy as np


def java_double_to_string(x):
    """java.lang.Double.toString: shortest digits that round-trip; plain decimal for 1e-3 <= |x| < 1e7, else d.dddE[-]n;
    always at least one digit after the point.  (Java renders at least two digits and, when one would do, picks the closest
    two-digit decimal -- this differs from the shortest form only for subnormals such as Double.MIN_VALUE = "4.9E-324".)"""
    x = float(x)
    if x != x:
        return "NaN"
    if x in (float("inf"), float("-inf")):
        return "Infinity" if x > 0 else "-Infinity"
    if x == 0.0:
        return "-0.0" if str(x).startswith("-") else "0.0"
    sign, digits, exp = Decimal(repr(x)).as_tuple()
    digits = list(digits)
    while len(digits) > 1 and digits[-1] == 0:          # strip trailing zeros of the digit string
        digits.pop(); exp += 1
    n = len(digits)
    e10 = exp + n - 1                                   # decimal exponent of the first digit
    ds = "".join(map(str, digits))
    if -3 <= e10 < 7:
        if e10 < 0:
            body = "0." + "0" * (-e10 - 1) + ds
        elif n <= e10 + 1:
            body = ds + "0" * (e10 + 1 - n) + ".0"
        else:
            body = ds[:e10 + 1] + "." + ds[e10 + 1:]
    else:
        body = ds[0] + "." + (ds[1:] or "0") + "E" + str(e10)
    return ("-" if sign else "") + body


def java_number_format(x, max_fraction_digits=5):
    """NumberFormat.getInstance() of the en-US default locale with setMaximumFractionDigits(5) (M:221-222): HALF_EVEN on the
    exact binary value, grouping commas, no trailing zeros, no forced fraction."""
    x = float(x)
    if x != x:
        return "NaN"                                    # CLDR symbols (JDK >= 9); JDK 8 prints U+FFFD
    if x in (float("inf"), float("-inf")):
        return ("-" if x < 0 else "") + "\u221e"
    q = Context(prec=400, rounding=ROUND_HALF_EVEN).quantize(Decimal(x), Decimal(1).scaleb(-max_fraction_digits))
    sign = "-" if math.copysign(1.0, x) < 0 else ""     # DecimalFormat keeps the sign of values that round to zero ("-0")
    s = format(abs(q), "f")
    ip, _, fp = s.partition(".")
    fp = fp.rstrip("0")
    groups = []
    while len(ip) > 3:
        groups.insert(0, ip[-3:]); ip = ip[:-3]
    groups.insert(0, ip)
    return sign + ",".join(groups) + ("." + fp if fp else "")


# ---- MALLET IDSorter order (cc.mallet.types.IDSorter.compareTo, bytecode): weight descending, then id DESCENDING ----------
def sorted_words(nwk_m, topic):
    """getSortedWords(m).get(topic), M:1792-1809: (type, count) of every type with a positive count, in TreeSet<IDSorter> order."""
    col = np.asarray(nwk_m)[:, topic]
    ids = np.nonzero(col > 0)[0]
    order = np.lexsort((-ids, -col[ids].astype(np.int64)))      # primary: count desc; secondary: id desc
    return [(int(ids[i]), int(col[ids[i]])) for i in order]


def top_words(nwk_m, num_words, lookup=str):
    """getTopWords(numWords, modality), M:1819-1845: per topic the first min(numWords, #positive) entries of sorted_words."""
    K = np.asarray(nwk_m).shape[1]
    return [[lookup(t) for t, _ in sorted_words(nwk_m, k)[:num_words]] for k in range(K)]


def display_top_words(nwk, alpha, lookups, num_words, using_new_lines=False):
    """displayTopWords(numWords, numLabels, usingNewLines), M:1851-1888.  Quirk kept: the loop starts at word = 1 and runs while
    word < numWords, so numWords - 1 words are shown per topic and view."""
    M, K = len(nwk), np.asarray(nwk[0]).shape[1]
    out = io.StringIO()
    for topic in range(K):
        for m in range(M):
            sw = sorted_words(nwk[m], topic)[:max(0, num_words - 1)]
            head = f"{topic}\t{java_number_format(alpha[m][topic])}"
            if using_new_lines:
                out.write(head + "\n")
                for t, c in sw:
                    out.write(f"{lookups[m](t)}\t{java_number_format(float(c))}\n")
            else:
                out.write(head + "\t")
                for t, _ in sw:
                    out.write(f"{lookups[m](t)}; ")
        out.write("\n")
    return out.getvalue()


def write_type_topic_counts(out, nwk, lookups):
    """printTypeTopicCounts, M:2076-2102: `type word 0:c0 1:c1 ...` -- every topic, zero counts included."""
    for m, tab in enumerate(nwk):
        tab = np.asarray(tab)
        for w in range(tab.shape[0]):
            out.write(f"{w} {lookups[m](w)}" + "".join(f" {t}:{int(c)}" for t, c in enumerate(tab[w])) + "\n")


def write_topic_word_weights(out, nwk, beta, lookups):
    """printTopicWordWeights, M:2113-2129: topic TAB word TAB (beta[m] + count) as Double.toString."""
    K = np.asarray(nwk[0]).shape[1]
    for topic in range(K):
        for m, tab in enumerate(nwk):
            col = np.asarray(tab)[:, topic]
            for w in range(len(col)):
                out.write(f"{topic}\t{lookups[m](w)}\t{java_double_to_string(float(beta[m]) + float(col[w]))}\n")


# ---- printState / its reader ------------------------------------------------------------------------------------------------
def write_state(out, views, zs, present, lookups, gamma, alpha, beta, sources=None):
    """printState(PrintStream), M:3276-3320.  views[m] = (doc_off, word_id), zs[m] = assignments in CSR order; present[m][d]
    tells whether document d owns an Assignments[m] object (the reference dereferences it unconditionally and would throw on a
    document that lacks the view, Q17 -- such views are skipped here).  `sources` (optional, per document) replaces "NA"."""
    M, K = len(views), np.asarray(alpha).shape[1] - 1
    out.write("#doc source pos typeindex type topic\n")
    out.write("#alpha : ")
    for m in range(M):
        out.write(f"modality:{m}\n")
        for t in range(K):
            out.write(java_double_to_string(float(gamma[m]) * float(alpha[m][t])) + " ")
    out.write("\n")
    out.write("#beta[0] : " + java_double_to_string(float(beta[0])) + "\n")
    D = len(views[0][0]) - 1
    for d in range(D):
        src = "NA" if sources is None or sources[d] is None else str(sources[d])
        for m in range(M):
            if present is not None and not present[m][d]:
                continue
            off, words = views[m]
            b, e = int(off[d]), int(off[d + 1])
            for pi in range(e - b):
                w = int(words[b + pi])
                out.write(f"{d} {src} {pi} {w} {lookups[m](w)} {int(zs[m][b + pi])}\n")


def write_state_gz(path, *args, **kw):
    """printState(File), M:3269-3274: the same text through a GZIPOutputStream."""
    with gzip.open(path, "wt", encoding="utf-8", newline="") as f:
        write_state(f, *args, **kw)


def read_state(inp, views, present=None):
    """Reads a printState file back into per-view assignment arrays for a corpus already held as CSR views (the file does not
    name the modality of a line: lines come per document, view after view, `pos` restarting at 0 -- the corpus tells how many
    lines each view of each document owns).  Checks doc / pos / typeindex of every line.  Returns (zs, header) where header
    holds the gamma*alpha rows and beta[0] found in the '#' lines."""
    if isinstance(inp, (str, bytes)):
        with gzip.open(inp, "rt", encoding="utf-8") as f:
            return read_state(f, views, present)
    M = len(views)
    zs = [np.full(len(v[1]), -1, dtype=np.int32) for v in views]
    header = {"alpha_lines": [], "beta0": None}
    D = len(views[0][0]) - 1
    d, m, pi = 0, 0, 0

    def advance():
        nonlocal d, m, pi
        while d < D:
            if m >= M:
                d += 1; m = 0; pi = 0
                continue
            absent = present is not None and not present[m][d]
            if absent or pi >= int(views[m][0][d + 1] - views[m][0][d]):
                m += 1; pi = 0
                continue
            return True
        return False

    in_header = True
    for line in inp:
        line = line.rstrip("\n")
        if in_header:                               # everything up to and including the "#beta[0] : x" line
            if line.startswith("#beta[0] : "):
                header["beta0"] = float(line[len("#beta[0] : "):])
                in_header = False
            elif not line.startswith("#doc "):
                header["alpha_lines"].append(line)
            continue
        if not line:
            continue
        f = line.split(" ")
        if len(f) < 6:
            raise ValueError(f"malformed state line: {line!r}")
        if not advance():
            raise ValueError("state file holds more token lines than the corpus")
        # `source` and the word itself may hold blanks (side-view labels such as "Deep Learning"), so the line is not split
        # blindly: doc is the first field, topic the last, and the (pos, typeindex) pair expected at this corpus position must
        # appear as two consecutive fields in between
        b = int(views[m][0][d])
        want = (str(pi), str(int(views[m][1][b + pi])))
        ok = f[0] == str(d) and any((f[i], f[i + 1]) == want for i in range(2, len(f) - 2))
        if not ok:
            raise ValueError(f"state line {line!r} does not match corpus position doc {d} view {m} pos {pi}")
        topic = int(f[-1])
        zs[m][b + pi] = topic
        pi += 1
    if advance():
        raise ValueError("state file ended before the corpus did")
    return zs, header


def topic_probabilities(topics, num_topics, gamma_m, alpha_m):
    """getTopicProbabilities(LabelSequence topics, byte modality), M:2148-2171: counts of the document's current assignments plus
    gamma[m] * alpha[m][t], normalised over the K topics (unassigned tokens, -1, are skipped: they hold no topic)."""
    import numpy as np
    t = np.asarray(topics, dtype=np.int64)
    d = np.bincount(t[t >= 0], minlength=num_topics).astype(np.float64)[:num_topics]
    d += float(gamma_m) * np.asarray(alpha_m, dtype=np.float64)[:num_topics]
    return d / d.sum()

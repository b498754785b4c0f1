"""ECPE text loader — SURVEY.md §8(f).4: `read_ECPE_data` of the trainers (drl_classifier_ec_mmd_final_mul.py:631-731,
identical in the hsic / vi / gan files), rebuilt for current pandas so that the unchanged trainers run end to end.

On-disk format (domains/README.md; e.g. domains/THUCTC_multiple/society_num.txt:1-12), per document:

    <doc_id> <n_clauses>
    (emo_clause, cause_clause), (emo_clause, cause_clause) ...      <- positive pairs
    <id>,<emotion_id>,<keyword>,<clause text>                       <- n_clauses lines; emotion_id 6 = no emotion

Output contract (same as the reference): `(df, docs_pair_size, num_unpred_emotions)` with `df` columns
`pair` ("<emotion clause>[SEP]<cause clause>", blanks removed, clause text = the 4th comma field only), `label`
(1 = true emotion-cause pair, 0 = sampled / enumerated negative) and `emotion` (the emotion id of the emotion clause);
one entry of `docs_pair_size` per document.  Training mode draws `random.sample(neg_pairs, k)` once per document in
file order — the same consumption of the `random` module's stream as the reference, so a seeded run yields the same
negatives.  What changed: rows are collected in a list and the frame is built once (the reference's per-row
`DataFrame.append` is quadratic and no longer exists in pandas >= 2), the pair line is parsed with
`ast.literal_eval` instead of `eval`, the file is closed, and `language` is an argument instead of the module-global
`opt.language`.  Host-only code: no GPU work here.
"""
from __future__ import annotations

import ast
import random
import re

import pandas as pd

_HEADER = re.compile(r"[0-9]{1,4}[\s][0-9]{1,2}")
NO_EMOTION = 6


def _clause_text(line: str) -> str:
    return line.strip().split(",")[3].replace(" ", "")


def _parse_pairs(line: str, language: str):
    line = line.strip()
    if language == "zh":                       # "(5,4), (7,6)": split on ", " like the reference (:647-648)
        return [tuple(ast.literal_eval(x)) for x in line.split(", ")]
    if language == "en":                       # "(2, 1), (4, 4),": one literal list (:650-652)
        return [tuple(p) for p in ast.literal_eval("[" + line + "]")]
    raise ValueError(f"language must be 'zh' or 'en', got {language!r}")


def read_ECPE_data(file_path, test: bool = False, language: str = "zh", rng=random):
    """See the module docstring.  `rng` needs `.sample(population, k)`; the default is the `random` module, as in
    the reference (seeded there with `random.seed(42)`, :27)."""
    rows = []
    docs_pair_size = []
    num_unpred_emotions = 0
    with open(file_path, encoding="utf8") as data_file:
        while True:
            line = data_file.readline()
            if not line:
                break
            if not _HEADER.search(line):
                continue
            doc_len = int(line.strip().split(" ")[1])
            pos_pairs = _parse_pairs(data_file.readline(), language)
            sentence_list, pred_emotions, sen_emo = [], [], {}
            for _ in range(doc_len):
                sentence = data_file.readline()
                sentence_list.append(sentence)
                fields = sentence.strip().split(",")
                sen_id, sen_emotion = int(fields[0]), int(fields[1])
                if sen_emotion != NO_EMOTION:
                    sen_emo[sen_id] = sen_emotion
                    pred_emotions.append(sen_id)

            if test:
                # keep only the true pairs whose emotion clause was predicted to carry an emotion (:668-682)
                keep, previous = [], -1
                for i, (e, _) in enumerate(pos_pairs):
                    if e not in pred_emotions and e != previous:
                        num_unpred_emotions += 1
                    elif e == previous:
                        keep.append(i)
                    else:
                        keep.append(i)
                        pred_emotions.remove(e)
                        previous = e
                pos_pairs = [pos_pairs[i] for i in keep]
            emotions = list(dict.fromkeys(e for e, _ in pos_pairs))
            causes = [c for _, c in pos_pairs]

            all_ids = range(1, doc_len + 1)
            neg_pairs = [(e, c) for e in emotions for c in all_ids if c not in causes]
            if not test:
                neg_pairs = rng.sample(neg_pairs, min(len(pos_pairs), len(neg_pairs)))
            else:
                neg_pairs += [(e, c) for e in pred_emotions for c in all_ids]

            for label, pairs in ((1, pos_pairs), (0, neg_pairs)):
                for e, c in pairs:
                    rows.append({"pair": _clause_text(sentence_list[e - 1]) + "[SEP]" + _clause_text(sentence_list[c - 1]),
                                 "label": label, "emotion": sen_emo[e]})
            docs_pair_size.append(len(pos_pairs) + len(neg_pairs))
    df = pd.DataFrame(rows, columns=["pair", "label", "emotion"])
    return df, docs_pair_size, num_unpred_emotions

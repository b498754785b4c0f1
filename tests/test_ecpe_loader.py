"""CPU: the ECPE text loader (carel_vae_b200/ecpe.py) against what the reference's own
`read_ECPE_data` (drl_classifier_ec_mmd_final_mul.py:631-731) returned on the same fixtures and `random` seeds
(tests/golden/ecpe_expected.json, written by oracle/make_golden_ecpe.py in the build container)."""
import json
import random
from pathlib import Path

import pytest

from carel_vae_b200 import ecpe

GOLDEN = Path(__file__).parent / "golden"
EXPECTED = json.loads((GOLDEN / "ecpe_expected.json").read_text(encoding="utf8"))
LANG = {"zh": "zh", "en": "en", "zh_unpred": "zh"}


@pytest.mark.parametrize("key", sorted(EXPECTED))
def test_matches_reference_reader(key):
    name, test, seed = key.split("/")
    test, seed = bool(int(test.split("=")[1])), int(seed.split("=")[1])
    want = EXPECTED[key]
    random.seed(seed)
    if "raises" in want:
        with pytest.raises(KeyError):
            ecpe.read_ECPE_data(GOLDEN / f"ecpe_{name}.txt", test=test, language=LANG[name])
        return
    df, sizes, unpred = ecpe.read_ECPE_data(GOLDEN / f"ecpe_{name}.txt", test=test, language=LANG[name])
    assert list(df.columns) == ["pair", "label", "emotion"]
    assert list(df["pair"]) == want["pair"]
    assert [int(x) for x in df["label"]] == want["label"]
    assert [int(x) for x in df["emotion"]] == want["emotion"]
    assert sizes == want["docs_pair_size"] and sum(sizes) == len(df)
    assert unpred == want["num_unpred_emotions"]


def test_training_negatives_follow_the_random_stream():
    """Different seeds give different negatives (the zh fixture has a document with 3 true pairs and 4 candidates);
    an explicit generator can replace the module-level stream."""
    a = EXPECTED["zh/test=0/seed=42"]["pair"]
    b = EXPECTED["zh/test=0/seed=7"]["pair"]
    assert a != b and sorted(EXPECTED["zh/test=0/seed=42"]["label"]) == sorted(EXPECTED["zh/test=0/seed=7"]["label"])
    df1, _, _ = ecpe.read_ECPE_data(GOLDEN / "ecpe_zh.txt", language="zh", rng=random.Random(42))
    assert list(df1["pair"]) == a


def test_clause_text_is_the_fourth_comma_field_only():
    """The reference takes `.split(',')[3]`: a clause that itself contains a comma is cut there (kept)."""
    df, _, _ = ecpe.read_ECPE_data(GOLDEN / "ecpe_zh.txt", test=True, language="zh")
    assert any(p.endswith("[SEP]家人也不理解") for p in df["pair"])


def test_bad_language_and_empty_file(tmp_path):
    with pytest.raises(ValueError):
        ecpe.read_ECPE_data(GOLDEN / "ecpe_zh.txt", language="fr")
    empty = tmp_path / "empty.txt"
    empty.write_text("")
    df, sizes, unpred = ecpe.read_ECPE_data(empty)
    assert len(df) == 0 and list(df.columns) == ["pair", "label", "emotion"] and sizes == [] and unpred == 0

"""TEST INFRASTRUCTURE — not product code.

Golden vectors for the ECPE text loader: runs the reference's OWN `read_ECPE_data`
(drl_classifier_ec_mmd_final_mul.py:631-731, AST-extracted — the script itself cannot be imported) on two
small hand-written fixtures and stores what it returns.  The reference builds its frame with `DataFrame.append`, which
pandas >= 2 no longer has; the shim below restores exactly that call (append one dict, ignore_index) and nothing else.

    python -m oracle.make_golden_ecpe        # build container only (/root/reference is not on the GPU box)

writes tests/golden/ecpe_{zh,en,zh_unpred}.txt (fixtures) and tests/golden/ecpe_expected.json.
"""
from __future__ import annotations

import ast
import json
import random
import re
import warnings
from pathlib import Path
from types import SimpleNamespace

import pandas as pd

REF = Path("/root/reference/drl_classifier_ec_mmd_final_mul.py")
OUT = Path(__file__).resolve().parent.parent / "tests" / "golden"

ZH = """\
10 6
 (5,4)
1,6,6,为了 坚持 到 妻子 入户
2,6,6,老 吴 已经 用尽 了 全力
3,6,6,老 吴 因为 其 抗癌 事迹 广为人知
4,6,4,如果 老 吴 长时间 不 更新 微博
5,4,担心,就 会 有 网友 担心 老 吴 是不是 去 了
6,6,6,2013 年底
11 5
 (2,1), (2,3), (4,4)
1,6,6,他 丢 了 工作
2,3,难过,心里 很 难过
3,6,6,家人 也 不 理解, 还 责怪 他
4,0,高兴,后来 找到 新 工作 很 高兴
5,2,6,只是 有些 累
12 2
 (1,1), (1,2)
1,5,惊讶,没 想到 会 这样
2,6,6,事情 来得 太 突然
"""

# the emotion clause of the true pair carries no emotion label: KeyError in training mode, counted (and dropped) in
# test mode — both are the reference's behaviour
ZH_UNPRED = ZH + """\
13 4
 (3,2)
1,1,生气,他 很 生气
2,6,6,因为 被 骗 了
3,6,6,于是 报 了 警
4,6,6,警察 来 了
"""

EN = """\
28 2
 (2, 1),
1,6,happiness,Think about surfing on the front edge of a wave
2,0,exhilarating,It 's really exhilarating .
118 6
 (4, 4), (4, 5), (6, 3),
1,6,null,We had a brave crew to pull us
2,6,null,but
3,6,null,pull as they would
4,2,fear,the current was so strong that we feared
5,6,null,if we persisted , we should be drawn away
6,4,sad,so we were sad
"""


def _append(self, row, ignore_index=False):
    return pd.concat([self, pd.DataFrame([row])], ignore_index=ignore_index)


def reference_reader(language: str):
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", SyntaxWarning)
        tree = ast.parse(REF.read_text())
    fn = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "read_ECPE_data"]
    ns = {"pd": pd, "re": re, "random": random, "opt": SimpleNamespace(language=language)}
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", SyntaxWarning)
        exec(compile(ast.Module(body=fn, type_ignores=[]), REF.name, "exec"), ns)
    return ns["read_ECPE_data"]


def main():
    OUT.mkdir(parents=True, exist_ok=True)
    (OUT / "ecpe_zh.txt").write_text(ZH, encoding="utf8")
    (OUT / "ecpe_en.txt").write_text(EN, encoding="utf8")
    (OUT / "ecpe_zh_unpred.txt").write_text(ZH_UNPRED, encoding="utf8")
    pd.DataFrame.append = _append
    expected = {}
    try:
        for name, language in (("zh", "zh"), ("en", "en"), ("zh_unpred", "zh")):
            read = reference_reader(language)
            for test in (False, True):
                for seed in (42, 7):
                    random.seed(seed)
                    try:
                        df, sizes, unpred = read(str(OUT / f"ecpe_{name}.txt"), test=test)
                        rec = {"pair": [str(x) for x in df["pair"]], "label": [int(x) for x in df["label"]],
                               "emotion": [int(x) for x in df["emotion"]], "docs_pair_size": [int(x) for x in sizes],
                               "num_unpred_emotions": int(unpred)}
                    except Exception as exc:                        # the reference's own failure modes are part of the contract
                        rec = {"raises": type(exc).__name__}
                    expected[f"{name}/test={int(test)}/seed={seed}"] = rec
    finally:
        del pd.DataFrame.append
    (OUT / "ecpe_expected.json").write_text(json.dumps(expected, ensure_ascii=False, indent=1), encoding="utf8")
    for k, v in expected.items():
        print(k, v.get("raises") or (len(v["pair"]), v["docs_pair_size"], v["num_unpred_emotions"]))


if __name__ == "__main__":
    main()

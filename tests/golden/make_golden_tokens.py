"""Generate tests/golden/tokens_*.json: the reference's class prompts tokenised by the reference's own tokenizer.

    python tests/golden/make_golden_tokens.py          (needs /root/reference: xlm-roberta-base/ tokenizer files, data/texts/*.json)

The GPU box has neither the tokenizer files nor the class-text lists, so the facade tests replay these fixtures through a stub
tokenizer (tests/util.FixtureTokenizer) that returns exactly what `AutoTokenizer(text=..., return_tensors="pt", padding=True)`
returned here (mm_backbone.py:382-383).  Prompts follow infer_wedetect.py:160-167: the first caption of every class, then ' '.
"""
import json
import os

from transformers import AutoTokenizer

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    tok = AutoTokenizer.from_pretrained(os.path.join(REF, "xlm-roberta-base"))
    for name, fn in (("coco_zh", "coco_zh_class_texts.json"), ("lvis_v1_zh", "lvis_v1_zh_class_texts.json")):
        classes = json.load(open(os.path.join(REF, "data", "texts", fn)))
        texts = [c[0] for c in classes] + [" "]
        enc = tok(text=texts, return_tensors="pt", padding=True)
        out = dict(source=f"data/texts/{fn} (first caption per class) + [' ']; tokenizer xlm-roberta-base/", texts=texts,
                   input_ids=enc["input_ids"].tolist(), attention_mask=enc["attention_mask"].tolist())
        with open(os.path.join(HERE, f"tokens_{name}.json"), "w") as f:
            json.dump(out, f, ensure_ascii=False)
        print(name, tuple(enc["input_ids"].shape))


if __name__ == "__main__":
    main()

"""Regenerates tests/golden/param_names.json from the reference's two parameter-name dumps
(train_svd_lora.txt = frozen names, train_svd_lora_train.txt = trainable names; written by reference
train_models/train_svd_lora.py:1249-1259).  Run in the dev container only (/root/reference is not on the GPU box)."""
import hashlib, json, pathlib
ref = pathlib.Path("/root/reference")
frozen = (ref / "train_svd_lora.txt").read_text().split()
train = (ref / "train_svd_lora_train.txt").read_text().split()
out = {"source": "reference train_svd_lora.txt / train_svd_lora_train.txt",
       "sha256": hashlib.sha256("\n".join(frozen + train).encode()).hexdigest(),
       "frozen": frozen, "trainable": train}
pathlib.Path(__file__).with_name("param_names.json").write_text(json.dumps(out, indent=0))
print(len(frozen), len(train))

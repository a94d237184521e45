"""Configuration tree with the field names federated_main.py / the trainer read (federated_main.py:60-153,
Dassl/dassl/config/defaults.py, configs/trainers/GLP_OT/vit_b16_oph.yaml).  A plain attribute namespace: yacs is
not needed on the hot path."""
from __future__ import annotations

import copy


class CfgNode(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v

    def clone(self):
        return copy.deepcopy(self)

    def merge_from_dict(self, d):
        for k, v in d.items():
            if isinstance(v, dict) and isinstance(self.get(k), CfgNode):
                self[k].merge_from_dict(v)
            else:
                self[k] = v
        return self


def _node(**kw):
    return CfgNode({k: (_node(**v) if isinstance(v, dict) else v) for k, v in kw.items()})


ATTRIBUTE_GROUPS = {   # GLP_OT_SVLoRA.retrieval_attributes (trainers/GLP_OT_SVLoRA.py:775-790)
    "FairFedMed": {"race": ["Asian", "Black", "White"], "language": ["English", "Spanish", "Others"],
                   "ethnicity": ["Non-hispanic", "Hispanic"], "gender": ["Male", "Female"]},
    "FedChexMimic": {"race": ["White", "Asian", "Black"], "gender": ["Male", "Female"], "age": ["0-60", "60+"]},
}


def get_cfg_default() -> CfgNode:
    """Defaults = the hyper-parameters of record (scripts/fairfedlora_fairfedmed.sh)."""
    return _node(
        SEED=1,
        OUTPUT_DIR="output",
        INPUT=dict(SIZE=(224, 224), PIXEL_MEAN=[0.48145466, 0.4578275, 0.40821073],
                   PIXEL_STD=[0.26862954, 0.26130258, 0.27577711], NO_TRANSFORM=True),
        DATASET=dict(NAME="FairFedMed", MODALITY_TYPE="slo_fundus", DIM_PER_3D_SLICE=8, USERS=3, ATTRIBUTES=["race"],
                     ATTRIBUTE_TYPE="race", NUM_TRAIN_PER_CLIENT=64, NUM_TEST_PER_CLIENT=64, SYNTHETIC=True),
        DATALOADER=dict(TRAIN_X=dict(BATCH_SIZE=32), TEST=dict(BATCH_SIZE=100), NUM_WORKERS=0),
        MODEL=dict(BACKBONE=dict(NAME="ViT-B/16"), INIT_WEIGHTS=""),
        OPTIM=dict(NAME="sgd", LR=1e-3, MOMENTUM=0.9, WEIGHT_DECAY=5e-4, MAX_EPOCH=1, LR_SCHEDULER="single_step",
                   STEPSIZE=(200,), GAMMA=0.1, ROUND=50),
        TRAINER=dict(NAME="GLP_OT_SVLoRA", LAMBDA_FAIRNESS=0.0,
                     GLP_OT=dict(N_CTX=4, CSC=False, CTX_INIT="", PREC="bf16", CLASS_TOKEN_POSITION="end", N=2,
                                 AVG_N=1, THRESH=1e-3, EPS=0.1, OT="None", TOP_PERCENT=0.8, MAX_ITER=100),
                     GLP_OT_LORA=dict(RANK=12, ALPHA=2.0, TYPE="FairLoRA", GLOBAL_S=False, LOCAL_S=False,
                                      DISABLE_ATTR=False, UNFREEZE_IMAGE_ENCODER=True)),
        TEST=dict(NO_TEST=False, PER_CLASS_RESULT=False),
        MODEL_ARCH=dict(VISION_LAYERS=12, VISION_WIDTH=768, PATCH=16, EMBED=512, TEXT_WIDTH=512, TEXT_LAYERS=12,
                        TEXT_HEADS=8, CONTEXT=77),
    )

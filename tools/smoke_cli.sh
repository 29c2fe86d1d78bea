#!/bin/bash
# End-to-end check of tools/train_net.py with reference-style KEY VALUE overrides (no dataset, synthetic weights):
# the ALDI++ flags on ResNet-50-FPN and on a small ConvNeXt-FPN, three iterations each on 128x160 images.
set -e
COMMON=(MODEL.ROI_HEADS.NUM_CLASSES 8 MODEL.RPN.PRE_NMS_TOPK_TRAIN 2000 MODEL.RPN.PRE_NMS_TOPK_TEST 1000
        MODEL.RPN.POST_NMS_TOPK_TRAIN 1000 MODEL.RPN.POST_NMS_TOPK_TEST 1000 SOLVER.IMS_PER_BATCH 4 SOLVER.IMS_PER_GPU 2
        SOLVER.AMP.ENABLED True SOLVER.WARMUP_ITERS 2 EMA.ENABLED True DOMAIN_ADAPT.TEACHER.ENABLED True
        DOMAIN_ADAPT.DISTILL.ROIH_CLS_ENABLED True DOMAIN_ADAPT.DISTILL.OBJ_ENABLED True
        DOMAIN_ADAPT.DISTILL.ROIH_REG_ENABLED True DOMAIN_ADAPT.DISTILL.RPN_REG_ENABLED True
        DOMAIN_ADAPT.DISTILL.HARD_ROIH_CLS_ENABLED False DATASETS.BATCH_CONTENTS "('labeled_strong','unlabeled_strong')"
        DATASETS.BATCH_RATIOS "(1,1)" MODEL.WEIGHTS models/missing.pkl)
python tools/train_net.py --iters 3 --image-size 128 160 SOLVER.BASE_LR 0.0005 "${COMMON[@]}" 2>&1 | tail -2
python tools/train_net.py --iters 3 --image-size 128 160 MODEL.BACKBONE.NAME build_convnext_fpn_backbone \
    MODEL.CONVNEXT.DEPTHS "[1,1,2,1]" MODEL.CONVNEXT.DIMS "[64,128,192,256]" SOLVER.OPTIMIZER ADAMW SOLVER.BASE_LR 0.0001 \
    "${COMMON[@]}" 2>&1 | tail -2

from oracle import torch_oracle as O


def non_max_suppression(prediction, conf_thres=0.25, iou_thres=0.45, classes=None, agnostic=False, multi_label=False,
                        labels=(), max_det=300, nc=0, max_time_img=0.05, max_nms=30000, max_wh=7680, in_place=True,
                        rotated=False):
    return O.non_max_suppression(prediction, conf_thres, iou_thres, classes, agnostic, multi_label, max_det, nc,
                                 max_nms, max_wh, in_place)

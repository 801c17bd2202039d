"""CPU restatement (numpy + OpenCV, as mmcv's image helpers call it) of the per-branch input pipeline
after RandomCrop / RandomFlip -- TEST INFRASTRUCTURE ONLY (tests/, pinned against the unmodified
``mmseg/datasets/pipelines/transforms.py`` by oracle/make_golden_pipeline.py).

  photometric_distortion   transforms.py:1197-1272 (convert / brightness / contrast / saturation / hue;
                           mmcv.bgr2hsv / hsv2bgr = cv2.cvtColor(COLOR_BGR2HSV / COLOR_HSV2BGR))
  imnormalize              mmcv.image.photometric.imnormalize_ as called at transforms.py:597-598
  pad + format             transforms.py:511-536 (mmcv.impad, pad_val 0 / seg_pad_val 255),
                           formatting.py:213-224 (HWC -> CHW, labels -> int64 [1, H, W])
"""
import cv2
import numpy as np


def convert(img, alpha=1, beta=0):
    img = img.astype(np.float32) * alpha + beta
    img = np.clip(img, 0, 255)
    return img.astype(np.uint8)


def photometric_distortion(img, params):
    """``params`` = (do_b, beta, mode, do_c, alpha_c, do_s, alpha_s, do_h, dh) as drawn by
    ``s4former_b200.datasets.draw_pmd_params`` (the reference draws them inline, same order)."""
    do_b, beta, mode, do_c, alpha_c, do_s, alpha_s, do_h, dh = params
    if do_b:
        img = convert(img, beta=beta)
    if mode == 1 and do_c:
        img = convert(img, alpha=alpha_c)
    if do_s:
        img = cv2.cvtColor(img, cv2.COLOR_BGR2HSV)
        img[:, :, 1] = convert(img[:, :, 1], alpha=alpha_s)
        img = cv2.cvtColor(img, cv2.COLOR_HSV2BGR)
    if do_h:
        img = cv2.cvtColor(img, cv2.COLOR_BGR2HSV)
        img[:, :, 0] = (img[:, :, 0].astype(int) + dh) % 180
        img = cv2.cvtColor(img, cv2.COLOR_HSV2BGR)
    if mode == 0 and do_c:
        img = convert(img, alpha=alpha_c)
    return img


def imnormalize(img, mean, std, to_rgb=True):
    img = img.copy().astype(np.float32)
    mean = np.float64(np.asarray(mean, dtype=np.float32).reshape(1, -1))
    stdinv = 1 / np.float64(np.asarray(std, dtype=np.float32).reshape(1, -1))
    if to_rgb:
        cv2.cvtColor(img, cv2.COLOR_BGR2RGB, img)
    cv2.subtract(img, mean, img)
    cv2.multiply(img, stdinv, img)
    return img


def branch(img_u8, label_u8, params, crop_size, mean, std, to_rgb=True, seg_pad_val=255):
    """-> (img [3, H, W] float32, gt [1, H, W] int64, distorted uint8 image [h, w, 3])."""
    d = photometric_distortion(img_u8.copy(), params)
    x = imnormalize(d, mean, std, to_rgb)
    H, W = crop_size
    pad = np.zeros((H, W, 3), dtype=np.float32)
    pad[:x.shape[0], :x.shape[1]] = x
    gt = np.full((H, W), seg_pad_val, dtype=np.uint8)
    if label_u8 is not None:
        gt[:label_u8.shape[0], :label_u8.shape[1]] = label_u8
    return np.ascontiguousarray(pad.transpose(2, 0, 1)), gt[None].astype(np.int64), d

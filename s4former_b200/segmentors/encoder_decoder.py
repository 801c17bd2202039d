"""``EncoderDecoder`` with the S4Former semi-supervised train step -- B200-native mirror of
``mmseg/models/segmentors/encoder_decoder.py`` (reference file:line cited per method).

The Python control flow of ``forward_train`` (:386-514) and ``foward_unsup_train`` (:516-687)
is kept (same kwargs, same loss-dict keys, same host RNG call order); the arithmetic is
dispatched to the CUDA library:

  EMA update (:1044-1066)                    -> one multi-tensor kernel          (ops.ema_update)
  teacher pseudo labels (:875-904, :541-542) -> one fused kernel incl. the patch
  + patch unconfidence (:547-555)               unconfidence                     (ops.pseudo_label)
  masked CE + NCR (:906-954)                 -> one fused fwd/bwd kernel         (ops.CeNcrFn)
  CutMix / PatchShuffle (:633-638)           -> two gather kernels               (utils.generate_unsup_data)

Options outside the shipped configs (unimatch, ClassMix, CutOut, fdrop, soft labels, neck,
projection heads, sup-branch NCR, ...) raise ``NotImplementedError`` at construction.
"""
import numpy as np
import torch
import torch.nn as nn

from .. import builder, ops
from ..builder import SEGMENTORS
from ..utils.generate_unsup_data import (draw_patchmix_perms, generate_cutout_box, generate_unsup_cutmix_data,
                                          generate_unsup_patchmix_data)
from ..utils.structual_utils import add_prefix, dict_split, weighted_loss
from .base import BaseSegmentor


@SEGMENTORS.register_module()
class EncoderDecoder(BaseSegmentor):
    def __init__(self,
                 backbone, decode_head, neck=None, auxiliary_head=None, projection_head=None,
                 backbone_ema=None, decode_head_ema=None, neck_ema=None, auxiliary_head_ema=None,
                 projection_head_ema=None, backbone_pretrain=None, pretrained=None, train_cfg=None,
                 test_cfg=None, init_cfg=None,
                 ema=False, sup_ema=False, ema_momentum=0.999, attn_frozen=False, attn_frozen_rate=0.0,
                 momentum_backbone=None, momentum_head=None, momentum_head_dropout=0.0,
                 momentum_head_exp=0.0, momentum_exp=0.0, ema_test=False,
                 sup_ClassMix=False, sup_cutmix=False,
                 unsup_weight=2.0, unsup_confidence=0.75, unsup_soft=False, unsup_temperature=1.0,
                 iter_unsup_start=0,
                 strong_aug_prob=0.5, cutout_area=2, use_CutMix=False, use_CutOut=False,
                 use_ClassMix=False, mix_with_labeled=False, patchwise=False,
                 use_PatchShuffle=False, PatchMix_N=8, patchmix_ratio=0.5, patchsize=16,
                 use_PatchShuffle_w_Classmix=False, use_PatchShuffle_w_Cutmix=False,
                 no_pos_embed=False, avg_pos_emd=False, duplicate_pos_emd=False,
                 adaptive_attn_mask=False, attn_mask_weight=50, attn_mask_seperate_head=False,
                 attn_mask_w_fdrop=False,
                 negative_class_ranking=False, negative_class_ranking_mode='sup_only',
                 use_fdrop=False, unimatch=False, fdrop_loss_weight=0.5, use_cutmix_adaptive=False):
        super().__init__(init_cfg)
        unsupported = dict(
            neck=neck, projection_head=projection_head, neck_ema=neck_ema,
            auxiliary_head_ema=auxiliary_head_ema, projection_head_ema=projection_head_ema,
            backbone_pretrain=backbone_pretrain, sup_ema=sup_ema, attn_frozen=attn_frozen,
            momentum_head_dropout=momentum_head_dropout, momentum_head_exp=momentum_head_exp,
            momentum_exp=momentum_exp, sup_ClassMix=sup_ClassMix, sup_cutmix=sup_cutmix,
            unsup_soft=unsup_soft, use_CutOut=use_CutOut, use_ClassMix=use_ClassMix,
            mix_with_labeled=mix_with_labeled, patchwise=patchwise, use_PatchShuffle=use_PatchShuffle,
            use_PatchShuffle_w_Classmix=use_PatchShuffle_w_Classmix, no_pos_embed=no_pos_embed,
            avg_pos_emd=avg_pos_emd, duplicate_pos_emd=duplicate_pos_emd,
            attn_mask_w_fdrop=attn_mask_w_fdrop, use_fdrop=use_fdrop, unimatch=unimatch,
            use_cutmix_adaptive=use_cutmix_adaptive)
        bad = [k for k, v in unsupported.items() if v]
        if bad or unsup_temperature != 1.0:
            raise NotImplementedError(
                f'options outside the S4Former train-step scope (SURVEY.md section 8): {bad}'
                + (' unsup_temperature' if unsup_temperature != 1.0 else ''))
        if negative_class_ranking and negative_class_ranking_mode != 'unsup_only':
            raise NotImplementedError("only negative_class_ranking_mode='unsup_only' (the shipped config)")
        if pretrained is not None:
            assert backbone.get('pretrained') is None, 'both backbone and segmentor set pretrained weight'
            backbone = dict(backbone, pretrained=pretrained)
        self.backbone = builder.build_backbone(backbone)
        self._init_decode_head(decode_head)
        self._init_auxiliary_head(auxiliary_head)
        self.train_cfg, self.test_cfg = train_cfg, test_cfg
        self.unsup_weight, self.unsup_confidence = unsup_weight, unsup_confidence
        self.unsup_soft, self.unsup_temperature = unsup_soft, unsup_temperature
        self.iter_unsup_start = iter_unsup_start
        self.ema, self.momentum = ema, ema_momentum
        self.momentum_backbone = momentum_backbone if momentum_backbone is not None else ema_momentum
        self.momentum_head = momentum_head if momentum_head is not None else ema_momentum
        self.ema_test = ema_test
        self.use_CutMix, self.cutout_area, self.strong_aug_prob = use_CutMix, cutout_area, strong_aug_prob
        self.PatchMix_N, self.patchmix_ratio, self.patchsize = PatchMix_N, patchmix_ratio, patchsize
        self.use_PatchShuffle_w_Cutmix = use_PatchShuffle_w_Cutmix
        self.negative_class_ranking = negative_class_ranking
        self.negative_class_ranking_mode = negative_class_ranking_mode
        self.fdrop_loss_weight = fdrop_loss_weight
        self.use_fdrop = False
        self.with_auxiliary_head_ema = False
        if self.ema:
            self._init_ema_model(pretrained, backbone_ema, decode_head_ema)
        assert self.with_decode_head
        self.attn_mask_weight, self.adaptive_attn_mask = attn_mask_weight, adaptive_attn_mask
        self.attn_mask_seperate_head = attn_mask_seperate_head
        self._ema_tables = {}
        self._topk_override = None   # parity tests may inject the PASA top-k index set
        # Run the three student backbone passes of the S4Former step (labeled, PASA, CutMix +
        # PatchShuffle) as ONE batched pass (see _forward_train_batched).  False keeps the
        # reference's pass-by-pass order.
        self.batch_student_passes = True
        self._step_aug = None        # pre-drawn augmentation parameters of the current step (TrainStep)

    # ------------------------------------------------------------------ construction (:164-246)
    def _init_ema_model(self, pretrained, backbone_ema, decode_head_ema):
        if pretrained is not None:
            backbone_ema = dict(backbone_ema, pretrained=pretrained)
        self.backbone_ema = builder.build_backbone(backbone_ema)
        for param in self.backbone_ema.parameters():
            param.detach_()
        self.decode_head_ema = builder.build_head(decode_head_ema)
        for param in self.decode_head_ema.parameters():
            param.detach_()

    def _init_decode_head(self, decode_head):
        self.decode_head = builder.build_head(decode_head)
        self.align_corners = self.decode_head.align_corners
        self.num_classes = self.decode_head.num_classes

    def _init_auxiliary_head(self, auxiliary_head):
        if auxiliary_head is not None:
            if isinstance(auxiliary_head, list):
                self.auxiliary_head = nn.ModuleList([builder.build_head(c) for c in auxiliary_head])
            else:
                self.auxiliary_head = builder.build_head(auxiliary_head)

    def init_weights(self):
        for m in self.children():
            if hasattr(m, 'init_weights'):
                m.init_weights()
            elif isinstance(m, nn.ModuleList):
                for mm in m:
                    if hasattr(mm, 'init_weights'):
                        mm.init_weights()

    # ------------------------------------------------------------------ features (:248-270)
    def extract_feat(self, img, no_pos_embed=False, avg_pos_emd=False, duplicate_pos_emd=False,
                     use_fdrop=False, attn_mask=None, attn_mask_weight=5, adaptive_attn_mask=False):
        return self.backbone(img, no_pos_embed=no_pos_embed, avg_pos_emd=avg_pos_emd,
                             duplicate_pos_emd=duplicate_pos_emd, use_fdrop=use_fdrop,
                             attn_mask=attn_mask, attn_mask_weight=attn_mask_weight,
                             adaptive_attn_mask=adaptive_attn_mask, topk_idx=self._topk_override)

    def extract_feat_ema(self, img):
        return self.backbone_ema(img)

    # ------------------------------------------------------------------ inference (:270-308, :1068-1265)
    def encode_decode(self, img, img_metas, adaptive_attn_mask=False, return_last_feat=False):
        """(:270-297) backbone + decode head + resize to the input size.  ``adaptive_attn_mask`` has
        a default here: the shipped ``whole_inference`` / ``slide_inference`` call this method
        without it and raise ``TypeError`` (SURVEY.md hazard 5); same signature otherwise."""
        if return_last_feat:
            raise NotImplementedError('return_last_feat (feature visualisation) is outside the hot path')
        x = self.extract_feat(img)
        out = self._decode_head_forward_test(x, img_metas)
        return ops.resize_bilinear(out, img.shape[2:])

    def encode_decode_ema(self, img, img_metas):
        """(:298-308)"""
        x = self.extract_feat_ema(img)
        out = self._decode_head_forward_test_ema(x, img_metas)
        return ops.resize_bilinear(out, img.shape[2:])

    def _decode_head_forward_test(self, x, img_metas, return_last_feat=False):
        return self.decode_head.forward_test(x, img_metas, self.test_cfg, return_last_feat)

    def _decode_head_forward_test_ema(self, x, img_metas):
        return self.decode_head_ema.forward_test(x, img_metas, self.test_cfg)

    @staticmethod
    def _cfg_get(cfg, key):
        return cfg[key] if isinstance(cfg, dict) else getattr(cfg, key)

    def slide_inference(self, img, img_meta, rescale):
        """(:1068-1115) sliding window with overlap; window logits are accumulated on the device."""
        h_stride, w_stride = self._cfg_get(self.test_cfg, 'stride')
        h_crop, w_crop = self._cfg_get(self.test_cfg, 'crop_size')
        batch_size, _, h_img, w_img = img.size()
        h_grids = max(h_img - h_crop + h_stride - 1, 0) // h_stride + 1
        w_grids = max(w_img - w_crop + w_stride - 1, 0) // w_stride + 1
        preds = torch.zeros((batch_size, self.num_classes, h_img, w_img), dtype=torch.float32, device=img.device)
        count_mat = torch.zeros((batch_size, 1, h_img, w_img), dtype=torch.float32, device=img.device)
        for h_idx in range(h_grids):
            for w_idx in range(w_grids):
                y1 = h_idx * h_stride
                x1 = w_idx * w_stride
                y2 = min(y1 + h_crop, h_img)
                x2 = min(x1 + w_crop, w_img)
                y1 = max(y2 - h_crop, 0)
                x1 = max(x2 - w_crop, 0)
                crop_img = img[:, :, y1:y2, x1:x2].contiguous()
                if self.ema_test:
                    crop_seg_logit = self.encode_decode_ema(crop_img, img_meta)
                else:
                    crop_seg_logit = self.encode_decode(crop_img, img_meta)
                ops.accumulate_crop(preds, count_mat, crop_seg_logit, y1, x1)
        preds = preds / count_mat          # (every pixel is covered: the grid spans the image)
        if rescale:
            resize_shape = img_meta[0]['img_shape'][:2]
            preds = preds[:, :, :resize_shape[0], :resize_shape[1]]
            preds = ops.resize_bilinear(preds, img_meta[0]['ori_shape'][:2])
        return preds

    def whole_inference(self, img, img_meta, rescale, tde=False, return_last_feat=False, use_attn_mask=None):
        """(:1117-1174)"""
        if tde or return_last_feat or use_attn_mask is not None:
            raise NotImplementedError('tde / return_last_feat / use_attn_mask are analysis paths')
        if self.ema_test:
            seg_logit = self.encode_decode_ema(img, img_meta)
        else:
            seg_logit = self.encode_decode(img, img_meta)
        if rescale:
            resize_shape = img_meta[0]['img_shape'][:2]
            seg_logit = seg_logit[:, :, :resize_shape[0], :resize_shape[1]]
            seg_logit = ops.resize_bilinear(seg_logit, img_meta[0]['ori_shape'][:2])
        return seg_logit

    def inference(self, img, img_meta, rescale, tde=False, return_last_feat=False, use_attn_mask=None,
                  want_pred=False):
        """(:1176-1212) softmax over classes of the whole / slide logits, flipped back when the test
        pipeline flipped the image.  ``want_pred=True`` also returns the arg-max map computed in the
        same kernel (``simple_test`` needs nothing else)."""
        mode = self._cfg_get(self.test_cfg, 'mode')
        assert mode in ['slide', 'whole']
        ori_shape = img_meta[0]['ori_shape']
        assert all(_['ori_shape'] == ori_shape for _ in img_meta)
        with torch.no_grad():
            if mode == 'slide':
                seg_logit = self.slide_inference(img, img_meta, rescale)
            else:
                seg_logit = self.whole_inference(img, img_meta, rescale, tde, return_last_feat, use_attn_mask)
            flip = img_meta[0].get('flip', False)
            direction = None
            if flip:
                direction = img_meta[0]['flip_direction']
                assert direction in ['horizontal', 'vertical']
            output, pred = ops.softmax_argmax(seg_logit, want_prob=True, want_pred=want_pred, flip=direction)
        return (output, pred) if want_pred else output

    def simple_test(self, img, img_meta, rescale=True, tde=False, return_last_feat=False):
        """(:1214-1232) -> list of per-image int64 label maps (numpy, like the reference)."""
        _, seg_pred = self.inference(img, img_meta, rescale, tde, return_last_feat, want_pred=True)
        return list(seg_pred.cpu().numpy())

    def simple_test_device(self, img, img_meta, rescale=True):
        """``simple_test`` without the device->host copy: [B, H, W] int64 on the device, ready for
        ``core.evaluation.intersect_and_union``."""
        return self.inference(img, img_meta, rescale, want_pred=True)[1]

    def aug_test(self, imgs, img_metas, rescale=True):
        """(:1257-1274) mean of the per-augmentation probabilities, then arg-max."""
        assert rescale
        seg_logit = self.inference(imgs[0], img_metas[0], rescale)
        for i in range(1, len(imgs)):
            seg_logit = seg_logit + self.inference(imgs[i], img_metas[i], rescale)
        seg_logit = seg_logit / len(imgs)
        seg_pred = seg_logit.argmax(dim=1)
        return list(seg_pred.cpu().numpy())

    def _decode_head_forward_train(self, x, img_metas, gt_semantic_seg):
        loss_decode = self.decode_head.forward_train(x, img_metas, gt_semantic_seg, self.train_cfg)
        return add_prefix(loss_decode, 'decode')

    def _auxiliary_head_forward_train(self, x, img_metas, gt_semantic_seg):
        losses = dict()
        if isinstance(self.auxiliary_head, nn.ModuleList):
            for idx, aux_head in enumerate(self.auxiliary_head):
                loss_aux = aux_head.forward_train(x, img_metas, gt_semantic_seg, self.train_cfg)
                losses.update(add_prefix(loss_aux, f'aux_{idx}'))
        else:
            loss_aux = self.auxiliary_head.forward_train(x, img_metas, gt_semantic_seg, self.train_cfg)
            losses.update(add_prefix(loss_aux, 'aux'))
        return losses

    # ------------------------------------------------------------------ forward_train (:386-514)
    def forward_train(self, img, img_metas, **kwargs):
        current_iter = kwargs.pop('iter')
        self.current_iter = current_iter
        kwargs.update({'img': img})
        kwargs.update({'img_metas': img_metas})
        kwargs.update({'tag': [meta['tag'] for meta in img_metas]})
        data_groups = dict_split(kwargs, 'tag')
        for _, v in data_groups.items():
            v.pop('tag')

        self.losses = dict()
        if self.ema:
            # ``_ema_fused_primed``: the optimizer applied this EMA to every parameter in the same
            # sweep as the previous SGD step (optim.FusedSGD.attach_ema) -- only the BatchNorm
            # running statistics (buffers, not optimizer state) are left to do here.
            params_done = bool(getattr(self, '_ema_fused_primed', False))
            with torch.no_grad():
                self.update_ema_variables(self.backbone, self.backbone_ema, self.momentum_backbone,
                                          buffers_only=params_done)
                self.update_ema_variables(self.decode_head, self.decode_head_ema, self.momentum_head,
                                          buffers_only=params_done)

        if self._can_batch_student_passes(data_groups, current_iter):
            self.losses.update(self._forward_train_batched(data_groups))
            return self.losses

        sup_imgs = sup_gts = None
        if 'sup' in data_groups:
            sup_imgs = data_groups['sup']['img']
            sup_gts = data_groups['sup']['gt_semantic_seg']
            labeled_features = self.extract_feat(sup_imgs)
            loss_decode_sup = self._decode_head_forward_train(
                labeled_features, data_groups['sup']['img_metas'], sup_gts)
            if self.with_auxiliary_head:
                loss_aux = self._auxiliary_head_forward_train(
                    labeled_features, data_groups['sup']['img_metas'], sup_gts)
                self.losses.update(loss_aux)
            self.losses.update(loss_decode_sup)

        if ('unsup_student' in data_groups) and self.unsup_weight != 0:
            unsup_loss = weighted_loss(
                self.foward_unsup_train(data_groups['unsup_teacher'], data_groups['unsup_student'],
                                        sup_imgs, sup_gts),
                weight=self.unsup_weight)
            if self.iter_unsup_start != 0:
                if current_iter > self.iter_unsup_start:
                    self.losses.update(unsup_loss)
            else:
                self.losses.update(unsup_loss)
        return self.losses

    # ------------------------------------------------------------------ batched student passes
    def _can_batch_student_passes(self, data_groups, current_iter):
        return (self.batch_student_passes and self.ema and self.attn_mask_seperate_head
                and self.use_PatchShuffle_w_Cutmix and not self.use_CutMix
                and 'sup' in data_groups and 'unsup_student' in data_groups
                and 'unsup_teacher' in data_groups and self.unsup_weight != 0
                and (self.iter_unsup_start == 0 or current_iter > self.iter_unsup_start)
                and data_groups['sup']['img'].shape[1:] == data_groups['unsup_student']['img'].shape[1:])

    @staticmethod
    def _feature_group(feats, b0, nb):
        """Images [b0, b0+nb) of a batched backbone output, as the per-pass tuple the heads take."""
        outs = []
        for f in feats:
            x2d, _, L = f._s4_tokens[:3]
            g = f[b0:b0 + nb]
            g._s4_tokens = (x2d, nb, L, b0)
            g._s4_hw = getattr(f, '_s4_hw', None)
            outs.append(g)
        return tuple(outs)

    def _forward_train_batched(self, data_groups):
        """The S4Former-full step (forward_train :426-512 + foward_unsup_train :516-687) with the
        three student backbone passes run as one batch [labeled | PASA | CutMix+PatchShuffle].

        The passes share weights and are independent of each other (the teacher is the only
        producer the unlabeled passes wait for), so batching them changes no value: same loss
        keys, same head / SyncBN invocation order (statistics stay per pass), same host RNG call
        order.  It cuts the kernel launches of the backbone by 3x and triples the rows of every
        GEMM (M = 24 x 1025 instead of 3 launches of 8 x 1025)."""
        sup = data_groups['sup']
        teacher_data, student_data = data_groups['unsup_teacher'], data_groups['unsup_student']
        sup_imgs, sup_gts = sup['img'], sup['gt_semantic_seg']
        loss_unsup = {}
        # ---- teacher (:519-542) ----
        tnames = [meta['filename'] for meta in teacher_data['img_metas']]
        snames = [meta['filename'] for meta in student_data['img_metas']]
        tidx = [tnames.index(name) for name in snames]
        with torch.no_grad():
            self.set_eval(self.ema)
            timg = teacher_data['img']
            if tidx != list(range(len(tidx))):
                timg = timg[self._index_tensor(tidx, timg.device)]
            tmetas = [teacher_data['img_metas'][idx] for idx in tidx]
            teacher_info = self.extract_teacher_info_ema(timg, tmetas)
            self.set_train(self.ema)
        student_info = self.extract_student_info(**student_data)
        attn_mask = self._patch_unconfidence(teacher_info, student_info)
        # the PASA pass sees the un-mixed images / labels / metas (:547-567)
        pasa_student = dict(student_info, img_metas=[dict(m) for m in student_info['img_metas']])
        pasa_teacher = dict(teacher_info)
        # ---- strong augmentation of the third pass (:633-638); host RNG order as the reference ----
        teacher_info, student_info = self._cutmix(teacher_info, student_info)
        student_info, teacher_info = self._patchmix(student_info, teacher_info)
        # ---- one backbone pass ----
        ns, nu = sup_imgs.shape[0], student_info['img'].shape[0]
        imgs = torch.cat([sup_imgs, pasa_student['img'], student_info['img']], 0)
        feats = self.backbone(imgs, attn_mask=attn_mask, attn_mask_weight=self.attn_mask_weight,
                              adaptive_attn_mask=self.adaptive_attn_mask, topk_idx=self._topk_override,
                              attn_mask_rows=(ns, ns + nu))
        # ---- heads and losses, in the reference's order ----
        losses = dict()
        labeled_features = self._feature_group(feats, 0, ns)
        loss_decode_sup = self._decode_head_forward_train(labeled_features, sup['img_metas'], sup_gts)
        if self.with_auxiliary_head:
            losses.update(self._auxiliary_head_forward_train(labeled_features, sup['img_metas'], sup_gts))
        losses.update(loss_decode_sup)
        # The branch weights (0.5 / fdrop_loss_weight, :567, :684-685) and unsup_weight
        # (weighted_loss, :488-512) are handed to the loss kernel, so the upstream gradient of every
        # term is 1 and the fused forward+gradient launch needs no rescale pass over dz.
        w = float(self.unsup_weight)
        pasa_student['backbone_feature'] = self._feature_group(feats, ns, nu)
        loss_unsup['loss_seg_unsup_attn_mask'] = \
            self.compute_pseudo_loss(pasa_student, pasa_teacher, want_ncr=False, ce_scale=0.5 * w)['loss_seg_unsup']
        student_info['backbone_feature'] = self._feature_group(feats, ns + nu, nu)
        mixed = self.compute_pseudo_loss(student_info, teacher_info, ce_scale=self.fdrop_loss_weight * w,
                                         ncr_scale=0.5 * w)
        if self.negative_class_ranking:
            loss_unsup['loss_ncr_unsup'] = mixed['loss_ncr_unsup']
        loss_unsup['loss_seg_unsup'] = mixed['loss_seg_unsup']
        losses.update(loss_unsup)
        return losses

    # ------------------------------------------------------------------ unsup branch (:516-687)
    def foward_unsup_train(self, teacher_data, student_data, sup_imgs, sup_gts):
        loss_unsup = {}
        tnames = [meta['filename'] for meta in teacher_data['img_metas']]
        snames = [meta['filename'] for meta in student_data['img_metas']]
        tidx = [tnames.index(name) for name in snames]

        with torch.no_grad():
            self.set_eval(self.ema)
            timg = teacher_data['img']
            if tidx != list(range(len(tidx))):
                timg = timg[self._index_tensor(tidx, timg.device)]
            tmetas = [teacher_data['img_metas'][idx] for idx in tidx]
            if not self.ema:
                teacher_info = self.extract_teacher_info(timg, tmetas)
            else:
                teacher_info = self.extract_teacher_info_ema(timg, tmetas)
            self.set_train(self.ema)
        # (:541-542) hard[conf==0] = 255 is already folded into the pseudo-label kernel

        student_info = self.extract_student_info(**student_data)

        if self.attn_mask_seperate_head:
            attn_mask = self._patch_unconfidence(teacher_info, student_info)
            unlabled_feat = self.extract_feat(
                student_info['img'], attn_mask=attn_mask, attn_mask_weight=self.attn_mask_weight,
                adaptive_attn_mask=self.adaptive_attn_mask)
            student_info['backbone_feature'] = unlabled_feat
            loss_unsup['loss_seg_unsup_attn_mask'] = \
                self.compute_pseudo_loss(student_info, teacher_info, want_ncr=False)['loss_seg_unsup'] * 0.5

        if self.use_CutMix:
            teacher_info, student_info = self._cutmix(teacher_info, student_info)

        if self.use_PatchShuffle_w_Cutmix:
            teacher_info, student_info = self._cutmix(teacher_info, student_info)
            student_info, teacher_info = self._patchmix(student_info, teacher_info)

        if not self.attn_mask_seperate_head:
            # Mean-Teacher config as shipped (:650-670): a PASA student pass whose features feed
            # no loss (hazard 4) -- executed for fidelity of the as-shipped step.
            # Its output reaches no loss, so it is run without autograd bookkeeping: same values,
            # no saved activations, and the pending-backward counters of the gradient reducer
            # (ops._pending_inc) are not armed for a backward that never comes.
            attn_mask = self._patch_unconfidence(teacher_info, student_info)
            with torch.no_grad():
                unlabled_feat = self.extract_feat(
                    student_info['img'], attn_mask=attn_mask, attn_mask_weight=self.attn_mask_weight,
                    adaptive_attn_mask=self.adaptive_attn_mask)
        else:
            unlabled_feat = self.extract_feat(student_info['img'])
        student_info['backbone_feature'] = unlabled_feat

        if self.use_fdrop or self.attn_mask_seperate_head:
            losses = self.compute_pseudo_loss(student_info, teacher_info)
            if self.negative_class_ranking:
                loss_unsup['loss_ncr_unsup'] = losses['loss_ncr_unsup'] * 0.5
            loss_unsup['loss_seg_unsup'] = losses['loss_seg_unsup'] * self.fdrop_loss_weight
        return loss_unsup

    # ------------------------------------------------------------------ strong augmentation
    def draw_aug_params(self, n_unsup, height, width):
        """ALL host RNG draws of one step, up front and in the reference's call order
        (encoder_decoder.py:604-607 CutMix, :633-638 CutMix + PatchShuffle; generate_unsup_data.py
        :7-26 three ``np.random.randint`` per box, :737-819 ``np.random.rand`` then
        ``torch.randperm`` per image).  A CutMix the coin rejects becomes all-zero boxes (the kernel
        then copies the images unchanged), so the step's device program is the same either way.
        Returns dict(cutmix=[int32 [n,4] ...], perms=int64 [n,num] or None) of CPU tensors."""
        out = dict(cutmix=[], perms=None)
        if n_unsup == 0 or self.unsup_weight == 0:
            return out

        def one_cutmix():
            if np.random.uniform(0, 1) < self.strong_aug_prob:
                boxes = [generate_cutout_box([height, width], ratio=self.cutout_area) for _ in range(n_unsup)]
            else:
                boxes = [(0, 0, 0, 0)] * n_unsup
            out['cutmix'].append(torch.tensor(boxes, dtype=torch.int32).reshape(n_unsup, 4))
        if self.use_CutMix:
            one_cutmix()
        if self.use_PatchShuffle_w_Cutmix:
            one_cutmix()
            out['perms'] = draw_patchmix_perms(n_unsup, height, width, self.patchmix_ratio, 16, self.PatchMix_N)
        return out

    def _cutmix(self, teacher_info, student_info):
        pre = self._step_aug
        if pre is not None:       # pre-drawn (resident) boxes: no host decision inside the step
            boxes = pre['cutmix'][pre['next_cutmix']]
            pre['next_cutmix'] += 1
            return generate_unsup_cutmix_data(teacher_info, student_info, ratio=self.cutout_area,
                                              patchwise=False, boxes=boxes)
        if np.random.uniform(0, 1) < self.strong_aug_prob:
            return generate_unsup_cutmix_data(teacher_info, student_info, ratio=self.cutout_area, patchwise=False)
        return teacher_info, student_info

    def _patchmix(self, student_info, teacher_info):
        pre = self._step_aug
        if pre is not None:
            return generate_unsup_patchmix_data(student_info, teacher_info, PatchMix_N=self.PatchMix_N,
                                                patchmix_ratio=self.patchmix_ratio, perms=pre['perms'],
                                                perms_dev=pre['perms_dev'])
        return generate_unsup_patchmix_data(student_info, teacher_info, PatchMix_N=self.PatchMix_N,
                                            patchmix_ratio=self.patchmix_ratio)

    def _index_tensor(self, idx, device):
        """Cached device copy of a host index list (no per-step host->device copy)."""
        cache = self.__dict__.setdefault('_s4_index_cache', {})
        key = (tuple(idx), str(device))
        t = cache.get(key)
        if t is None:
            t = torch.tensor(idx, device=device)
            cache[key] = t
        return t

    def _patch_unconfidence(self, teacher_info, student_info):
        """(:547-555) u = mean over each patch of (1 - conf); produced by the pseudo-label kernel."""
        if teacher_info['conf_mask'].shape[-1] != student_info['img'].shape[-1]:
            raise NotImplementedError('MiT-style 1/4-resolution teacher maps are a "next" row (SegFormer)')
        return teacher_info['patch_unconf']

    def extract_student_info(self, img, img_metas, gt_semantic_seg, **kwargs):
        """(:832-850)"""
        if 'valid' in img_metas[0]:
            raise NotImplementedError("'valid' masks are not produced by the shipped pipelines")
        return dict(img=img, img_metas=img_metas, gt_semantic_seg=gt_semantic_seg[0])

    def _teacher_outputs(self, seg_logits, unsup_confidence):
        thr = unsup_confidence if unsup_confidence is not None else self.unsup_confidence
        if thr is None or thr == 0:
            raise NotImplementedError('unsup_confidence=0 (no confidence mask) is not a shipped setting')
        hard, conf, u = ops.pseudo_label(seg_logits, thr, self.patchsize)
        return hard, conf, u

    def extract_teacher_info(self, img, img_metas, unsup_confidence=None):
        """(:852-873) teacher = student weights (ema=False)."""
        feat = self.extract_feat(img)
        seg_logits = self.decode_head.forward_get_logits(feat, self.train_cfg, img_metas)
        hard, conf, u = self._teacher_outputs(seg_logits, unsup_confidence)
        return dict(backbone_feature=feat, seg_logits=seg_logits, hard_seg_label=hard, conf_mask=conf,
                    patch_unconf=u, img_metas=img_metas)

    def extract_teacher_info_ema(self, img, img_metas, unsup_confidence=None):
        """(:875-904)"""
        feat = self.extract_feat_ema(img)
        seg_logits = self.decode_head_ema.forward_get_logits(feat, self.train_cfg, img_metas)
        hard, conf, u = self._teacher_outputs(seg_logits, unsup_confidence)
        return dict(backbone_feature=feat, seg_logits=seg_logits, hard_seg_label=hard, conf_mask=conf,
                    patch_unconf=u, img_metas=img_metas)

    def compute_pseudo_loss(self, student_info, teacher_info, want_ncr=True, ce_scale=1.0, ncr_scale=1.0):
        """(:906-954) masked CE (mean over ALL pixels) and, in 'unsup_only' mode, NCR.
        ``want_ncr=False`` skips the NCR value the reference computes and discards (:567);
        ``ce_scale`` / ``ncr_scale`` multiply the two terms inside the kernel."""
        loss_unsup = {}
        students_prediction = self.decode_head.forward_get_logits(
            student_info['backbone_feature'], self.train_cfg, student_info['img_metas'])
        ncr = self.negative_class_ranking and want_ncr and \
            self.negative_class_ranking_mode in ('unsup_only', 'both')
        zt = teacher_info['seg_logits'] if ncr else None
        loss_ce, loss_ncr, _ = ops.CeNcrFn.apply(students_prediction, zt, teacher_info['hard_seg_label'],
                                                 float(ce_scale), float(ncr_scale) if ncr else 0.0, 255)
        loss_unsup['loss_seg_unsup'] = loss_ce
        if ncr:
            loss_unsup['loss_ncr_unsup'] = loss_ncr
        return loss_unsup

    # ------------------------------------------------------------------ EMA (:1044-1066)
    def ema_pairs(self):
        """{id(student parameter): (teacher parameter, momentum)} in the reference's pairing order
        (zip of named_parameters(), encoder_decoder.py:1049-1051) for the fused SGD + EMA sweep."""
        out = {}
        if not self.ema:
            return out
        for model, ema_model, m in ((self.backbone, self.backbone_ema, self.momentum_backbone),
                                    (self.decode_head, self.decode_head_ema, self.momentum_head)):
            for (_, sp_), (_, tp) in zip(model.named_parameters(), ema_model.named_parameters()):
                out[id(sp_)] = (tp, float(m))
        return out

    def update_ema_variables(self, model, ema_model, momentum=0.999, dropout=0.0, attn_frozen=False,
                             buffers_only=False):
        if dropout != 0 or attn_frozen:
            raise NotImplementedError('EMA dropout / attn_frozen are not used by the shipped configs')
        key = (id(model), id(ema_model), bool(buffers_only))
        table = self._ema_tables.get(key)
        src = [] if buffers_only else [p for _, p in model.named_parameters()]
        dst = [] if buffers_only else [p for _, p in ema_model.named_parameters()]
        for (sn, sb), (_, tb) in zip(model.named_buffers(), ema_model.named_buffers()):
            if 'bn' in sn and 'num_batches_tracked' not in sn:
                src.append(sb)
                dst.append(tb)
        if not dst:
            return
        shadows = ops.shadow_list(dst)      # bf16 copies of the teacher weights, refreshed in the same pass
        ptr_key = tuple(t.data_ptr() for t in dst) + tuple(t.data_ptr() for t in src) + \
            tuple(0 if s is None else s.data_ptr() for s in shadows)
        if table is None or table.key != ptr_key:
            if not dst[0].is_cuda:
                raise RuntimeError('s4former_b200 needs CUDA tensors: there is no CPU fallback')
            table = ops.TensorTable([[t.data for t in dst], [t.data for t in src], shadows], dst[0].device)
            table.targets = dst
            self._ema_tables[key] = table
        ops.ema_update(table, momentum)
        for t in table.targets:
            ops.bump_generation(t)
        ops.mark_shadows_fresh(table.targets)

    def _set_mode(self, ema, training):
        """`.train(mode)` of the student (or EMA) sub-models.  The module list is cached and the
        flag written straight into each module's ``__dict__``: ``nn.Module.train`` re-walks ~400
        modules through ``nn.Module.__setattr__`` and was ~2 ms of host time per step for the two
        teacher toggles (reference encoder_decoder.py:524,539)."""
        cache = self.__dict__.setdefault('_s4_mode_lists', {})
        mods = cache.get(bool(ema))
        if mods is None:
            roots = [self.backbone_ema, self.decode_head_ema] if ema else \
                [self.backbone, self.decode_head] + ([self.auxiliary_head] if self.with_auxiliary_head else [])
            mods = [m for r in roots for m in r.modules()]
            # VisionTransformer.train keeps its LayerNorms in eval mode under norm_eval (vit.py:416-445)
            frozen = [m for r in roots if getattr(r, 'norm_eval', False)
                      for m in r.modules() if isinstance(m, torch.nn.LayerNorm)]
            mods = (mods, frozen)
            cache[bool(ema)] = mods
        for m in mods[0]:
            m.__dict__['training'] = training
        for m in mods[1]:
            m.__dict__['training'] = False

    def set_eval(self, ema=False):
        self._set_mode(ema, False)

    def set_train(self, ema=False):
        self._set_mode(ema, True)

"""Checkpoint naming / layout — mirrors reference utils/utils.py:110-167 for the DDPM archs.

Layout: ``torch.save({"opt": optimizer.state_dict(), "model": denoiser.state_dict()}, path)``
with ``path = cfg.DATA_FS.SAVE_DIR + cfg.MODEL.NAME.format(arch, EPOCHS, PAST_LEN, FUTURE_LEN,
epoch_tag, "NA")``; loaded with ``torch.load(..., weights_only=True)['model']``.
"""
import logging
import os

import torch


def get_backbone_cfg(cfg, arch):
    gen_model_key, backbone_key = arch.upper().split('-')
    return getattr(getattr(cfg.MODEL, gen_model_key), backbone_key)


def get_model_fullname(cfg, arch, epoch):
    total_epochs = get_backbone_cfg(cfg, arch).TRAIN.EPOCHS
    if arch in ("DDPM-UNet", "DDPM-DiT"):
        return cfg.DATA_FS.SAVE_DIR + cfg.MODEL.NAME.format(
            arch, total_epochs, cfg.DATASET.PAST_LEN, cfg.DATASET.FUTURE_LEN, epoch, "NA")
    if arch in ("FM-UNet", "FM-DiT"):
        return cfg.DATA_FS.SAVE_DIR + cfg.MODEL.NAME.format(
            arch, total_epochs, cfg.DATASET.PAST_LEN, cfg.DATASET.FUTURE_LEN, epoch, cfg.MODEL.FM.W_TYPE)
    logging.error("Architecture not supported.")
    return None


get_checkpoint_save_path = get_model_fullname


def save_checkpoint(optimizer, model, epoch, cfg, arch):
    path = get_checkpoint_save_path(cfg, arch, epoch)
    d = os.path.dirname(path)
    if d:
        os.makedirs(d, exist_ok=True)
    torch.save({"opt": optimizer.state_dict(), "model": model.state_dict()}, path)
    return path


def create_directory(directory):
    if directory and not os.path.exists(directory):
        os.makedirs(directory)

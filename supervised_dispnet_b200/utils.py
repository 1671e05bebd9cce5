"""Checkpoint I/O in the reference's on-disk format (reference: utils.py:79-93, written at train.py:375-386, read back at
train.py:273-281, 322-323): two files per save, `dispnet_<filename>` and `exp_pose_<filename>`, each a torch-pickled dict with
'epoch' and 'state_dict' (+ 'optimizer' for the disparity net); the best one is copied to `<prefix>_model_best.pth.tar`.
The drop-in models keep the reference's state_dict keys and shapes, so files written by either side load on the other."""
import os
import shutil

import torch


def save_checkpoint(save_path, dispnet_state, exp_pose_state, is_best, epoch, filename='checkpoint.pth.tar', record=False):
    save_path = str(save_path)
    file_prefixes = ['dispnet', 'exp_pose']
    states = [dispnet_state, exp_pose_state]
    for (prefix, state) in zip(file_prefixes, states):
        torch.save(state, os.path.join(save_path, '{}_{}'.format(prefix, filename)))
    if record:
        record_path = os.path.join(save_path, 'weights_{}'.format(epoch))
        os.makedirs(record_path, exist_ok=True)
        torch.save(dispnet_state, os.path.join(record_path, 'dispnet_{}'.format(filename)))
    if is_best:
        for prefix in file_prefixes:
            shutil.copyfile(os.path.join(save_path, '{}_{}'.format(prefix, filename)),
                            os.path.join(save_path, '{}_model_best.pth.tar'.format(prefix)))


def load_checkpoint(path, net, optimizer=None, strict=True, map_location='cpu'):
    """weights = torch.load(path); net.load_state_dict(weights['state_dict']) (train.py:280-281, :322-323)."""
    weights = torch.load(str(path), map_location=map_location, weights_only=False)
    net.load_state_dict(weights['state_dict'], strict=strict)
    if optimizer is not None and 'optimizer' in weights:
        optimizer.load_state_dict(weights['optimizer'])
    return weights.get('epoch')

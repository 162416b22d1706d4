"""Constants of the inference path (mirror of the reference's Config/config.py:11-70: same attribute names,
same values), with checkpoint/sample paths pointing into this repository's Resource/ tree."""
import os

import numpy as np
import torch

_root = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


class Config:
    frame_no = 20
    pc_no = 128
    batch_size = 20
    lower_pc_no = 64
    joint_num_all = 21
    joint_num_upper = 15
    joint_num_lower = 8
    num_action = 13
    IMU_used = True
    dataset_random_seed = 1

    device = "cuda:0" if torch.cuda.is_available() else "cpu"

    skeleton_all = np.asarray(
        [[20, 3], [3, 2], [2, 1], [2, 4], [2, 8], [4, 5], [5, 6], [6, 7], [8, 9], [9, 10], [10, 11],
         [1, 0], [0, 12], [0, 16], [12, 13], [13, 14], [14, 15], [16, 17], [17, 18], [18, 19]])
    skeleton_upper_body = skeleton_all[:14]
    skeleton_lower_body = skeleton_all[14:]

    kinect_upper_gragh = [(0, 12), (0, 13), (0, 1), (1, 2), (2, 3), (2, 4), (2, 8), (3, 14), (4, 5), (5, 6), (6, 7),
                          (8, 9), (9, 10), (10, 11)]
    kinect_joint_selection = [0, 1, 2, 3, 4, 5, 6, 7, 11, 12, 13, 14, 18, 19, 20, 21, 22, 23, 24, 25, 26]
    upper_joint_map = [0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 16, 20]
    lower_joint_map = [12, 13, 14, 15, 16, 17, 18, 19]

    model_IMU_path = os.path.join(_root, "Resource/Pretrained_model/IMU_Net/epoch173_batch20frame20lr3e-05.pth")
    model_upper_path = os.path.join(_root, "Resource/Pretrained_model/Upper_Net/epoch451_batch20frame20lr3e-05.pth")
    model_lower_path = os.path.join(_root, "Resource/Pretrained_model/Lower_Net/epoch161_batch20frame20lr0.0003.pth")
    # frozen batch tensors of Resource/Sample_data (built by the reference loader with np.random.seed(0);
    # oracle/make_golden.py --full-sample)
    sample_frozen_path = os.path.join(_root, "Resource/Sample_data_frozen/sample835_seed0.npz")
    # raw per-frame cache for the GPU snippet builder (scripts/pack_sample_data.py; not shipped -- 150 MB of .mat input)
    sample_packed_path = os.path.join(_root, "Resource/Sample_data_packed/raw.npz")
    sample_data_path = os.path.join(_root, "Resource/Sample_data")

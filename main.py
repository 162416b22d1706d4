"""`python main.py --infer` -- same command line as the reference's main.py:5-73 for the inference path.
The flag set is the reference's (`--network`, `--train`, `--infer`, `--vis`, `--colab`, `--epochs`, `--lr`, `--device`,
`--batch_size`, `--log_dir`, `--load_*_path`); training and visualisation are outside this repository's scope and
exit saying so.  `--imu_surrogate` and `--from_raw` are additions."""
import argparse
import sys


def main():
    p = argparse.ArgumentParser(description="mmEgo inference on B200")
    p.add_argument("--network", type=str, choices=["IMU_Net", "Upper_Net", "Lower_Net"],
                   help="(reference flag, main.py:8) network to train -- training is not implemented here")
    p.add_argument("--train", action="store_true", help="(reference flag, main.py:10) not implemented here")
    p.add_argument("--infer", action="store_true", help="evaluate on Resource/Sample_data (frozen tensors)")
    p.add_argument("--vis", action="store_true", help="(reference flag) visualisation -- not implemented here")
    p.add_argument("--epochs", type=int)
    p.add_argument("--lr", type=float)
    p.add_argument("--device", type=str, help="cuda device, e.g. cuda:0")
    p.add_argument("--batch_size", type=int, help="snippets per batch (the reference's eval loop hard-codes 1)")
    p.add_argument("--log_dir", type=int)
    p.add_argument("--load_IMU_path", type=str)
    p.add_argument("--load_Upper_path", type=str)
    p.add_argument("--load_Lower_path", type=str)
    p.add_argument("--colab", action="store_true")
    p.add_argument("--imu_surrogate", action="store_true",
                   help="feed IMU_Net's training targets as (R, t) (needed while the IMU checkpoint is missing)")
    p.add_argument("--from_raw", action="store_true",
                   help="build the batches on the GPU from the packed raw sensor cache (Resource/Sample_data_packed) "
                        "instead of reading the frozen tensors")
    a = p.parse_args()

    from mmego_b200.Config.config import Config
    if a.device:
        Config.device = a.device
    if a.batch_size:
        Config.batch_size = a.batch_size
    if a.load_IMU_path:
        Config.model_IMU_path = a.load_IMU_path
    if a.load_Upper_path:
        Config.model_upper_path = a.load_Upper_path
    if a.load_Lower_path:
        Config.model_lower_path = a.load_Lower_path
    if a.train or a.vis:
        sys.exit("only --infer is implemented: training (--train --network ...) and visualisation (--vis) are out of "
                 "scope for the B200 inference path")
    if not a.infer:
        p.print_help()
        return
    from mmego_b200.Processor.Test.Demo_test import MMEgo
    MMEgo(imu_surrogate=True if a.imu_surrogate else None, from_raw=a.from_raw).eval_model()


if __name__ == "__main__":
    main()

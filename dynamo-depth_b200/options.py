"""Command-line options.  The flag names, defaults and per-dataset fill-ins are the reference's public
API (options.py:4-303) and are kept identical so `python train.py ...` lines written for the
reference keep working; loss weights are still discovered through the `g_` prefix (Trainer.py:299).
"""
import argparse

# (flags, kwargs) -- table driven so the whole surface is visible at a glance
_ARGS = [
    # experiment
    (("--model_name", "-n"), dict(type=str, default="--", help="folder name of this run")),
    (("--log_dir",), dict(type=str, default="./logs")),
    (("--eval_dir",), dict(type=str, default="./outputs")),
    # system
    (("--cuda_ids",), dict(nargs="+", type=int, default=[0], help="one id per local rank")),
    (("--local_rank", "--local-rank"), dict(type=int, default=0, dest="local_rank")),
    (("--ddp",), dict(type=bool, default=False, help="set by train.py when LOCAL_WORLD_SIZE > 1")),
    (("--num_workers",), dict(type=int, default=2)),
    # dataset
    (("--dataset", "-d"), dict(type=str, choices=["kitti", "waymo", "nuscenes"], default="waymo")),
    (("--data_path",), dict(type=str, default=None)),
    (("--split",), dict(type=str, default=None)),
    (("--height",), dict(type=int, default=None)),
    (("--width",), dict(type=int, default=None)),
    (("--img_ext",), dict(type=str, choices=[".png", ".jpg"], default=".jpg")),
    (("--cam_name",), dict(type=str, default=None)),
    # loss weights (every g_* attribute becomes a loss term)
    (("--g_p_photo",), dict(type=float, default=1.0)),
    (("--g_d_smooth",), dict(type=float, default=1e-3)),
    (("--g_d_ground",), dict(type=float, default=0.1)),
    (("--g_c_smooth",), dict(type=float, default=1e-3)),
    (("--g_c_consistency",), dict(type=float, default=5.0)),
    (("--g_m_sparsity",), dict(type=float, default=0.04)),
    (("--g_m_smooth",), dict(type=float, default=0.1)),
    (("--weight_ramp",), dict(nargs="+", type=str, default=["g_c_smooth", "g_c_consistency", "g_m_sparsity", "g_m_smooth"])),
    (("--ramp_red",), dict(type=float, default=3)),
    (("--ssim_weight",), dict(type=float, default=0.85)),
    (("--mask_disp_thrd",), dict(type=float, default=0.03)),
    # training hyper-parameters
    (("--epoch_schedules",), dict(nargs="+", type=int, default=[1, 1, 5, 20], help="disp_init motion_init mask_init fine_tune")),
    (("--epoch-size",), dict(type=int, default=8000)),
    (("--batch_size", "-b"), dict(type=int, default=3)),
    (("--learning_rate",), dict(type=float, default=1e-4)),
    (("--scheduler_step_size",), dict(type=int, default=10)),
    # model
    (("--depth_model",), dict(type=str, choices=["monodepthv2", "litemono"], default="litemono")),
    (("--encoder_num_layers",), dict(type=int, default=18, choices=[18, 34, 50, 101, 152])),
    (("--weights_init",), dict(type=str, default="pretrained", choices=["pretrained", "scratch"])),
    (("--scales",), dict(nargs="+", type=int, default=None)),
    # training options
    (("--frame_ids",), dict(nargs="+", type=int, default=[0, -1, 1])),
    (("--min_depth",), dict(type=float, default=0.1)),
    (("--max_depth",), dict(type=float, default=100.0)),
    (("--train_img_type",), dict(type=str, choices=["original", "downsample"], default=None)),
    # ground plane (RANSAC)
    (("--gp_prior",), dict(type=float, default=0.4)),
    (("--gp_tol",), dict(type=float, default=0.005)),
    (("--gp_max_it",), dict(type=int, default=100)),
    (("--gp_np_per_it",), dict(type=int, default=5)),
    # loading / logging
    (("--load_ckpt", "-l"), dict(type=str, default="")),
    (("--log_frequency",), dict(type=int, default=100)),
    (("--no_train_vis",), dict(action="store_true")),
    (("--save_frequency",), dict(type=int, default=1)),
    (("--comment", "-c"), dict(type=str, default="")),
    (("--print_opt",), dict(type=bool, default=True)),
    # evaluation
    (("--eval_min_depth",), dict(type=float, default=1e-3)),
    (("--eval_max_depth",), dict(type=float, default=None)),
    (("--eval_img_bound",), dict(nargs="+", type=int, default=None)),
    (("--eval_img_ext",), dict(type=str, choices=[".png", ".jpg"], default=None)),
    (("--eval_img_type",), dict(type=str, choices=["original", "downsample"], default=None)),
    # B200 build only (not in the reference; defaults keep the reference's behaviour)
    (("--skip_unused_depth",), dict(action="store_true")),     # no depth passes on frames -1/+1 (no loss term reads them)
    (("--batch_pose_pairs",), dict(action="store_true")),      # both pose-encoder calls as one batch of 2B pairs (changes BN batch statistics)
    (("--encoder_tf32_linear",), dict(action="store_true")),   # Lite-Mono linear layers in single-pass TF32 (cuBLAS)
    # Lite-Mono linear layers: "tc3x" = tcgen05 3xTF32 kernel at fp32 accuracy (csrc/linear_tc.cu), "torch" = torch fp32 matmul
    (("--encoder_linear",), dict(type=str, default="tc3x", choices=["tc3x", "torch"])),
    # continue an interrupted run: a models/<phase>_<epoch> folder written by Trainer.save_model (weights + trainer_state.pth)
    (("--resume",), dict(type=str, default="")),
]

# values filled in when the corresponding option is left at None (options.py:274-301)
_PER_DATASET = {
    "split": {"waymo": "waymo", "nuscenes": "nuscenes", "kitti": "eigen_zhou"},
    "height": {"waymo": 320, "nuscenes": 288, "kitti": 192},
    "width": {"waymo": 480, "nuscenes": 512, "kitti": 640},
    "cam_name": {"waymo": "FRONT", "nuscenes": "FRONT", "kitti": "image_02"},
    "train_img_type": {"waymo": "downsample", "nuscenes": "downsample", "kitti": "downsample"},
    "eval_max_depth": {"waymo": 75, "nuscenes": 75, "kitti": 80},
    "eval_img_bound": {"waymo": [0, 1, 0, 1], "nuscenes": [0, 1, 0, 1],
                       "kitti": [0.40810811, 0.99189189, 0.03594771, 0.96405229]},
    "eval_img_ext": {"waymo": ".jpg", "nuscenes": ".jpg", "kitti": ".png"},
    "eval_img_type": {"waymo": "downsample", "nuscenes": "downsample", "kitti": "original"},
}

_DEFAULT_SCALES = {"monodepthv2": [0, 1, 2, 3], "litemono": [0, 1, 2]}


class DynamoOptions:
    def __init__(self):
        self.p = argparse.ArgumentParser(description="Dynamo options")
        for flags, kw in _ARGS:
            self.p.add_argument(*flags, **kw)

    def parse(self, **kwargs):
        self.opt = self.p.parse_args(**kwargs)
        if self.opt.scales is None:
            self.opt.scales = list(_DEFAULT_SCALES[self.opt.depth_model])
        if self.opt.data_path is None:
            self.opt.data_path = f"data_dir/{self.opt.dataset}/"
        for key, val in list(vars(self.opt).items()):
            if val is None:
                setattr(self.opt, key, _PER_DATASET[key][self.opt.dataset])
        return self.opt

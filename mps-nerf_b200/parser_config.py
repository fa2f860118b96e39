"""Flag system with the reference's surface (parser_config.py:3-113): same flags, defaults,
``--config`` files of ``key = value`` lines, CLI overrides the file, and argparse prefix
abbreviations (configs/h36m.txt relies on ``i_test`` -> ``--i_testset``).

configargparse is not a dependency: the small subset the reference uses is implemented here.
"""
import argparse
import sys


class ConfigArgumentParser(argparse.ArgumentParser):
    def __init__(self, *a, **k):
        super().__init__(*a, **k)
        self._config_dests = []

    def add_argument(self, *a, **k):
        is_cfg = k.pop("is_config_file", False)
        act = super().add_argument(*a, **k)
        if is_cfg:
            self._config_dests.append(act.option_strings[0])
        return act

    @staticmethod
    def _file_args(path):
        out = []
        with open(path) as fh:
            for line in fh:
                line = line.split("#")[0].strip()
                if not line or "=" not in line:
                    continue
                key, val = (s.strip() for s in line.split("=", 1))
                if val.lower() == "true":
                    out.append("--" + key)
                elif val.lower() == "false":
                    continue
                else:
                    out += ["--" + key, val]
        return out

    def parse_known_args(self, args=None, namespace=None):
        args = list(sys.argv[1:] if args is None else args)
        file_args = []
        for opt in self._config_dests:
            for i, a in enumerate(args):
                if a == opt and i + 1 < len(args):
                    file_args += self._file_args(args[i + 1])
                elif a.startswith(opt + "="):
                    file_args += self._file_args(a.split("=", 1)[1])
        return super().parse_known_args(file_args + args, namespace)   # later (CLI) values win


def config_parser():
    p = ConfigArgumentParser()
    A = p.add_argument
    A("--config", is_config_file=True, help="config file path")
    A("--expname", type=str, help="experiment name")
    A("--basedir", type=str, default="./logs/", help="where to store ckpts and logs")
    # training options
    A("--N_rand", type=int, default=1024 * 32, help="random rays per gradient step")
    A("--lrate", type=float, default=5e-4)
    A("--decay_steps", type=int, default=10000)
    A("--chunk", type=int, default=1024 * 64, help="rays processed in parallel")
    A("--netchunk_per_gpu", type=int, default=1024 * 64 * 64)
    A("--no_reload", action="store_true")
    A("--ft_path", type=str, default=None)
    # rendering options
    A("--N_samples", type=int, default=64, help="coarse samples per ray")
    A("--N_importance", type=int, default=0)
    A("--perturb", type=float, default=1.0)
    A("--use_viewdirs", action="store_true")
    A("--with_viewdirs", type=int, default=1)
    # dataset options
    A("--data_root", type=str, default="msra_h36m/S9/Posing")
    A("--data_set_type", type=str, default="multi_pair")
    A("--train_split", type=str, default="test")
    A("--test_split", type=str, default="test")
    A("--image_scaling", type=float, default=0.4)
    A("--model", type=str, default="correction_by_f3d")
    A("--N_iteration", type=int, default=48001)
    A("--white_bkgd", action="store_true")
    for name, d in (("use_os_env", 0), ("multi_person", 1), ("density_loss", 0), ("correction_loss", 0),
                    ("acc_loss", 1), ("T_loss", 1), ("smooth_loss", 1), ("consistency_loss", 0), ("half_acc", 0),
                    ("human_sample", 0), ("num_worker", 8), ("start", 0), ("interval", 10), ("poses_num", 100),
                    ("num_instance", 100), ("test_num_instance", 1), ("random_pair", 1), ("use_f2d", 0),
                    ("use_trans", 0), ("save_weights", 1), ("view_num", 3), ("border", 5), ("batch_size", 1),
                    ("local_rank", 0), ("ddp", 0), ("occupancy", 0), ("mean_shape", 1), ("correction_field", 0),
                    ("skinning_field", 0), ("smooth_interval", 4), ("append_rgb", 1), ("male", 0), ("new_mask", 0),
                    ("test_persons", 2), ("ani_nerf_ft", 0), ("i_print", 120), ("i_weights", 12000),
                    ("i_testset", 3000), ("smpl_shape_loss", 1)):
        A("--" + name, type=int, default=d)
    # extension (not in the reference): arithmetic of the MLP/transformer kernels
    A("--precision", type=str, default="bf16", choices=["bf16", "fp32"])
    return p


def print_args(args):
    print("--------args----------")
    for k, v in vars(args).items():
        print("%s: %s" % (k, v))
    print("--------args----------\n")

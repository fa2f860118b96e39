"""Model factory with the reference's name (model_selection.py:8-130)."""
from .lib.skinnning_batch import SKinningBatch


def return_model(global_args):
    if global_args.model != "skinning_batch":
        raise NotImplementedError(
            f"model '{global_args.model}': mpsnerf_b200 builds the hot path of the shipped configs "
            "(model = skinning_batch, configs/canonical_transformer.txt:5, configs/h36m.txt:30) only")
    return SKinningBatch(
        human_sample=global_args.human_sample, density_loss=global_args.density_loss,
        with_viewdirs=global_args.with_viewdirs, use_f2d=global_args.use_f2d, use_trans=global_args.use_trans,
        smooth_loss=global_args.smooth_loss, num_instances=global_args.num_instance,
        mean_shape=global_args.mean_shape, correction_field=global_args.correction_field,
        skinning_field=global_args.skinning_field, data_set_type=global_args.data_set_type,
        append_rgb=global_args.append_rgb, precision=getattr(global_args, "precision", None))

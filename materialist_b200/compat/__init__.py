"""Reference-side binding: stand-ins for the `mitsuba` and `drjit` modules that route the calls the reference's scripts make
(inverse_img_w_mi.py, render_final.py, trans_edit.py, myutils/misc.py) to the B200 operator, so that those scripts run unchanged:

    import materialist_b200.compat as compat
    compat.install()                      # sys.modules['mitsuba'], sys.modules['drjit'] = the stand-ins
    import inverse_img_w_mi               # the reference's own module, from its own tree

What is mapped (reference file:line -> here):
  mi.load_dict({...'type': 'scene'...})   inverse_img_w_mi.py:40-56, render_final.py:23-97   -> SceneSpec -> materialist_b200.Scene (traced PLY)
  mi.traverse(scene) / params[...] = T / params.update()   inverse_img_w_mi.py:61-64, :72-78, :216-220   -> ParamsProxy over scene.traverse
  mi.render(scene, params, spp=, seed=)   inverse_img_w_mi.py:65, :79; render_final.py:194, :227, :391   -> materialist_b200.render, with the
      tensors assigned through `params` that require grad attached as differentiable leaves (what dr.wrap_ad does in the reference)
  dr.wrap_ad(source='torch', target='drjit')   inverse_img_w_mi.py:59, :69   -> identity decorator (the operator is a torch.autograd.Function)
  mi.Bitmap / mi.util.write_bitmap / mi.TensorXf   misc.py:99-146, inverse_img_w_mi.py:641-743   -> native image readers / writers, torch tensors
  mi.register_bsdf / mi.set_variant / dr.set_flag / mi.OptixDenoiser   -> accepted and recorded (MatDiffBSDF / TransBSDF are built into the
      operator; the OptiX AI denoiser is out of scope: the stand-in returns its input)

Scenes are described lazily (SceneSpec) and built on first use, so the dictionary handling is testable without a GPU."""
import sys

from . import drjit_shim, mitsuba_shim


def install(force=True):
    """Register the stand-ins as `mitsuba` / `drjit`.  force=False keeps real modules if they are importable."""
    if not force:
        try:
            import mitsuba  # noqa: F401
            import drjit  # noqa: F401
            return False
        except Exception:
            pass
    sys.modules["mitsuba"] = mitsuba_shim.module()
    sys.modules["drjit"] = drjit_shim.module()
    return True

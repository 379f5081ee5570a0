"""Loading the reference's config files for the hot-path modules.

The reference reads `configs/3d-multi-resolution-rcnn.py` with `mmcv.Config.fromfile` (tools/train.py:46,
tools/test.py:107), which executes the Python file and exposes its module-level names.  mmcv is not a dependency
here; `load_config` does the same thing with `runpy` and returns a plain dict, and `build_from_config` constructs
the modules of the RoI hot path from the keys the reference's detector hands to them:
    model.rpn_head            -> RPNProposal3D          (configs/3d-multi-resolution-rcnn.py:16-27)
    model.bbox_roi_extractor  -> SingleRoIExtractor      (:38-45)
    model.mask_roi_extractor  -> SingleRoIExtractor      (:66-73)
    test_cfg.rpn / test_cfg.rcnn, train_cfg.rpn_proposal (:105-111, :132-143)
Everything else in the file (backbone, heads, datasets, schedules) belongs to the reference and is left alone.
"""
import runpy

from .builder import build_roi_extractor, build_rpn_proposal

HOT_PATH_KEYS = ('rpn_head', 'bbox_roi_extractor', 'mask_roi_extractor')


def load_config(path):
    """Execute a reference config file; returns {name: value} of its public module-level names."""
    ns = runpy.run_path(path)
    return {k: v for k, v in ns.items() if not k.startswith('_')}


def hot_path_config(cfg):
    """The sub-dicts of a loaded config that parameterise the RoI hot path (JSON-serialisable)."""
    model = cfg['model']
    out = {k: model[k] for k in HOT_PATH_KEYS if k in model}
    out['bbox_head_target_stds'] = model.get('bbox_head', {}).get('target_stds')
    out['test_cfg'] = cfg.get('test_cfg')
    train = cfg.get('train_cfg') or {}
    out['train_rpn_proposal'] = train.get('rpn_proposal')
    return out


def build_from_config(cfg):
    """cfg: load_config(path), or the dict hot_path_config() returns.  Returns (rpn, bbox_extractor, mask_extractor)."""
    model = cfg['model'] if 'model' in cfg else cfg
    rpn = build_rpn_proposal(model['rpn_head'])
    bbox_ex = build_roi_extractor(model['bbox_roi_extractor'])
    mask_ex = build_roi_extractor(model['mask_roi_extractor']) if model.get('mask_roi_extractor') else None
    return rpn, bbox_ex, mask_ex

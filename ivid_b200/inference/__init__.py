from .sample import sample_all, shard, build_modelviews
from .utils import parse_int_list, save_scene, load_scene, reorder

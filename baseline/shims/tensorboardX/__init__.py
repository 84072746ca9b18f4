"""Stand-in for tensorboardX (not installed in this image; harl/utils/configs_tools.py:71 imports SummaryWriter from it):
scalars are appended to <log_dir>/scalars.jsonl.  Used only when the unmodified HARL runner is driven by
dc_rl_b200.harl_runner; with the real package on the path this directory is never imported."""
import json
import os


class SummaryWriter:
    def __init__(self, log_dir="./runs", **kwargs):
        os.makedirs(log_dir, exist_ok=True)
        self._f = open(os.path.join(log_dir, "scalars.jsonl"), "a", encoding="utf-8")

    def add_scalar(self, tag, value, global_step=None, **kwargs):
        self._f.write(json.dumps({"tag": tag, "value": float(value), "step": global_step}) + "\n")

    def add_scalars(self, main_tag, tag_scalar_dict, global_step=None, **kwargs):
        for k, v in tag_scalar_dict.items():
            self.add_scalar("%s/%s" % (main_tag, k), v, global_step)

    def export_scalars_to_json(self, path):
        pass

    def flush(self):
        self._f.flush()

    def close(self):
        self._f.close()

"""Saved-configuration discovery (rl/utils/config.py:5-21): the single JSON ``*.conf`` of a save dir."""
import json
import os


def get_conf(save_dir):
    found = [fn for fn in os.listdir(save_dir) if fn.endswith('.conf')]
    if len(found) > 1:
        raise Exception('too many confs in directory')
    if not found:
        return None
    with open(os.path.join(save_dir, found[0])) as fin:
        return json.load(fin)

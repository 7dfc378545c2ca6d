"""Logging configuration (cf. the reference's test log4j.properties, LOG4J:1-13).

``configure_logging(path)`` applies a ``logging.config.dictConfig`` JSON file
(``configs/logging.json`` is the quiet test profile).  Importing the package
applies ``$GW2V_LOG_CONFIG`` automatically, so shard-server processes spawned
by ``fit`` inherit the same settings through the environment.
"""
from __future__ import annotations

import json
import logging
import logging.config
import os
from typing import Optional


def configure_logging(path: Optional[str] = None) -> bool:
    path = path or os.environ.get("GW2V_LOG_CONFIG")
    if not path:
        return False
    with open(path) as f:
        cfg = json.load(f)
    cfg.pop("_comment", None)
    logging.config.dictConfig(cfg)
    return True

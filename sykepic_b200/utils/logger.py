"""Logging setup of the reference (sykepic/utils/logger.py:14-34): `LOGLEVEL` env, same format."""

import logging
import os
from logging.config import dictConfig
from pathlib import Path

SETUP_RAN = False


def get_logger(name):
    global SETUP_RAN
    if not SETUP_RAN:
        setup()
        SETUP_RAN = True
    return logging.getLogger(name)


def setup(config_file=None):
    if config_file:
        import yaml

        with open(config_file) as fh:
            config = yaml.safe_load(fh.read())
        Path(config["handlers"]["file"]["filename"]).parent.mkdir(parents=True, exist_ok=True)
        dictConfig(config)
    else:
        logging.basicConfig(
            level=os.environ.get("LOGLEVEL", "INFO"),
            format="{asctime} - {name} - {levelname} - {message}",
            style="{",
        )

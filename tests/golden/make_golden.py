#!/usr/bin/env python3
"""Regenerates tests/golden/ristretto_snapshots.json from the reference's own snapshot fixtures.

Run in the build container only (it reads /root/reference, which does not exist on the GPU box):

    python tests/golden/make_golden.py

Source: /root/reference/tests/snapshots/snapshots__*-ristretto.snap, produced by
/root/reference/tests/snapshots.rs:31-189 from `ChaChaRng::seed_from_u64(12345)`.
The `.snap` files are `insta` YAML; values are base64url (unpadded) strings.  They are converted to
hex here so that the tests need no YAML/base64 handling; no value is altered.
"""
import base64
import json
import pathlib

import yaml

SNAP_DIR = pathlib.Path("/root/reference/tests/snapshots")
OUT = pathlib.Path(__file__).with_name("ristretto_snapshots.json")


def b64hex(s):
    pad = "=" * (-len(s) % 4)
    return base64.urlsafe_b64decode(s + pad).hex()


def convert(node):
    if isinstance(node, dict):
        return {k: convert(v) for k, v in node.items()}
    if isinstance(node, list):
        return [convert(v) for v in node]
    if isinstance(node, str):
        return b64hex(node)
    return node


def main():
    out = {"_source": "slowli/elastic-elgamal tests/snapshots/*-ristretto.snap (seed_from_u64(12345)), hex-encoded"}
    for path in sorted(SNAP_DIR.glob("snapshots__*-ristretto.snap")):
        name = path.name[len("snapshots__"):-len("-ristretto.snap")]
        text = path.read_text()
        # insta format: '---\n<metadata>\n---\n<body>'
        parts = text.split("\n---\n", 1)
        body = parts[1] if len(parts) == 2 else text
        doc = yaml.safe_load(body)
        out[name] = convert(doc)
    OUT.write_text(json.dumps(out, indent=1, sort_keys=True) + "\n")
    print("wrote", OUT, "with", len(out) - 1, "snapshots")


if __name__ == "__main__":
    main()

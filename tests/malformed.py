"""Malformed-input cases: one per bail!/anyhow! site of the reference's wire primitives
(ruhvro/src/fast_decode.rs:575,591,646,849,866,874,884,898,906,910)."""
import json

from oracle import pyoracle as po

FLAT = json.dumps({"type": "record", "name": "F", "fields": [
    {"name": "i", "type": "int"}, {"name": "b", "type": "boolean"}, {"name": "s", "type": "string"},
    {"name": "d", "type": "double"}, {"name": "f", "type": "float"},
    {"name": "n", "type": ["null", "long"]}, {"name": "u", "type": ["string", "int", "boolean"]},
    {"name": "e", "type": {"type": "enum", "name": "E", "symbols": ["A", "B"]}},
    {"name": "a", "type": {"type": "array", "items": "int"}}]})


def good_record(i=0):
    s = po.parse_schema(FLAT)
    return po.encode_datum(s, {"i": i, "b": True, "s": "hello", "d": 1.5, "f": 2.5, "n": (1, 7), "u": (1, 3), "e": 1, "a": [1, 2, 3]})


def cases():
    """(name, expected error code name, bad record bytes)"""
    z = po.zigzag_bytes
    head = z(1) + b"\x01" + z(5) + b"hello" + b"\0" * 8 + b"\0" * 4   # i, b, s, d, f
    out = [
        ("eof_in_varint", "eof", b"\x80"),
        ("eof_empty", "eof", b""),
        ("varint_too_long", "varint", b"\xff" * 10 + b"\x01"),
        ("bad_bool", "bool", z(1) + b"\x02"),
        ("neg_string_len", "neg_len", z(1) + b"\x01" + z(-3)),
        ("string_past_end", "eof", z(1) + b"\x01" + z(50) + b"abc"),
        ("eof_double", "eof", z(1) + b"\x01" + z(0) + b"\0\0\0"),
        ("eof_float", "eof", z(1) + b"\x01" + z(0) + b"\0" * 8 + b"\0\0"),
        ("bad_null_branch", "branch", head + z(2)),
        ("bad_union_branch", "branch", head + z(0) + z(3)),
        ("neg_union_branch", "branch", head + z(0) + z(-1)),
        ("bad_enum", "enum", head + z(0) + z(1) + z(0) + z(2)),
        ("neg_enum", "enum", head + z(0) + z(1) + z(0) + z(-1)),
        ("array_truncated", "eof", head + z(0) + z(1) + z(0) + z(0) + z(5) + z(1)),
        ("array_no_terminator", "eof", head + z(0) + z(1) + z(0) + z(0) + z(1) + z(1)),
    ]
    return out

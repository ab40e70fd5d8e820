"""ORACLE -- CPU restatement of the reference path.  Test infrastructure only: importable from
tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs, never from hd_yolo_b200/."""

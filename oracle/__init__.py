"""ORACLE: CPU restatement of the reference hot path. Test infrastructure only -- nothing under
strique_b200/ may import this package ("Oracle")."""

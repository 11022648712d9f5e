"""Name-only stand-ins so /root/reference/modules/controlresiduals_pipeline.py imports (annotators are off the hot path)."""


class _Detector:
    @classmethod
    def from_pretrained(cls, *a, **k):
        raise RuntimeError("annotators are out of scope (SURVEY.md §2 row 7)")


LineartDetector = LineartAnimeDetector = PidiNetDetector = OpenposeDetector = _Detector
MLSDdetector = NormalBaeDetector = HEDdetector = _Detector

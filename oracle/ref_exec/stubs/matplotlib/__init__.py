"""matplotlib stand-in: imported by run_video.py for rendering only, which is outside the path."""

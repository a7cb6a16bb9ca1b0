"""TEST INFRASTRUCTURE ONLY -- import-only stand-in for matplotlib (rad_search_env.py:6, 17-21; render() is out of scope)."""

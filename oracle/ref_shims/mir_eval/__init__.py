"""empty stand-in (sonification is off the hot path)"""

"""ColorType — mirror of pixo::ColorType (src/color.rs:8-48)."""
import enum


class ColorType(enum.IntEnum):
    Gray = 0
    GrayAlpha = 1
    Rgb = 2
    Rgba = 3

    def bytes_per_pixel(self) -> int:
        return (1, 2, 3, 4)[int(self)]

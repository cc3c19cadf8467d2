// C target needs one translation unit (the reference ships empty.c too: Sources/RendererTypes/empty.c).

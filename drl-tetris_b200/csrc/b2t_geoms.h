// Field sizes the library is instantiated for (H, W).  The reference takes any game_size; each size here costs one
// instantiation of every kernel, so the list is the BASELINE.json configurations, two small ones used by tests and 25x12 — the tallest field whose columns still fit
// 32-bit words (H + 7 <= 32) with a width above 10 (the garbage quirk of gamePlay.cpp:202 and the 256-entry lists).
// Add a line to support another size.
#pragma once
#define B2T_FOR_EACH_GEOMETRY(X) \
    X(22, 10)                    \
    X(16, 8)                     \
    X(40, 20)                    \
    X(10, 6)                     \
    X(12, 5)                     \
    X(25, 12)

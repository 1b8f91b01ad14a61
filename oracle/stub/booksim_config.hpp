// Stand-in for the Booksim2 fork's config header (third-party, network-fetched
// by the reference build, absent offline). Only the cycle-accurate timing model
// uses Booksim2; that model is out of scope and unpinned.
#pragma once
#include <string>
#include <vector>
class BookSimConfig {};

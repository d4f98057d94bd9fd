/* oracle/ref_shim: empty stand-in so that the reference's common.h (which only forward-uses OIIO::TextureSystem) compiles. TEST INFRASTRUCTURE. */
#pragma once
namespace OIIO { class TextureSystem; }

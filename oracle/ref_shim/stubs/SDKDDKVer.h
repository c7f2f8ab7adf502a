/* stub of the Windows SDK version header included by the reference's PSBA/targetver.h */
#pragma once

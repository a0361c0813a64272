// Oracle build shim: intentionally empty (nothing from this header is used on the sync path).
#ifndef COS_SHIM_BOOST_DATE_TIME_POSIX_TIME_POSIX_TIME_HPP_
#define COS_SHIM_BOOST_DATE_TIME_POSIX_TIME_POSIX_TIME_HPP_
#endif
